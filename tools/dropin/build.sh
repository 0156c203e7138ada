#!/bin/bash
# Drop-in proof: compiles the UNCHANGED reference sources -- the gtest suite stereoDNN/tests/tests_main.cpp, the four
# generated network builders stereoDNN/sample_app/*_net.cpp, sample_app/main.cpp and the TrailNet runtime of ros/packages/caffe_ros -- where they lie under /root/reference, against this repo's
# headers (include/NvInfer.h, redtail_tensorrt_plugins.h, internal_utils.h) and links them to libnvstereo_inference.so.
# Outputs go to dropin/_ref/ (git-ignored, travels to the GPU box like the other built binaries).  Only runs where
# /root/reference exists; nothing is copied from it.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF=/root/reference/stereoDNN
[ -d "$REF" ] || { echo "no reference checkout at $REF: skipping drop-in build"; exit 0; }
OUT="$ROOT/dropin/_ref"
mkdir -p "$OUT"
INC="-I$ROOT/include -I$ROOT/tools/dropin/include -I/usr/local/cuda/include"
LIB="-L$ROOT/redtail_b200/lib -lnvstereo_inference -lredtail_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,\$ORIGIN/../../redtail_b200/lib"
CXX="${CXX:-g++} -std=c++14 -O1 -w"
for n in nvsmall_1025x321 nvtiny_513x161 resnet18_1025x321 resnet18_2D_513x257; do
  $CXX $INC -c "$REF/sample_app/${n}_net.cpp" -o "$OUT/${n}_net.o"
done
$CXX $INC -c "$REF/tests/tests_main.cpp" -o "$OUT/tests_main.o"
$CXX $INC -c "$ROOT/tools/dropin/gtest_lite.cpp" -o "$OUT/gtest_lite.o"
$CXX -o "$OUT/nvstereo_tests" "$OUT/tests_main.o" "$OUT/gtest_lite.o" $LIB
# The builders are linked into a small driver (tools/dropin/net_driver.cpp) that calls create<Net>Network() exactly like
# sample_app/main.cpp does and runs the engine on raw .bin images.
$CXX $INC -I"$REF/sample_app" -c "$ROOT/tools/dropin/net_driver.cpp" -o "$OUT/net_driver.o"
$CXX -o "$OUT/nvstereo_net_driver" "$OUT/net_driver.o" "$OUT"/*_net.o $LIB
# Plans of the reference's ResNet networks at the resolution of BASELINE.json's configs, written host-only by the
# reference's own builders (net driver `dump` mode): what bench.py --config resnet18 / resnet18_2d runs.
mkdir -p "$OUT/plans"
W="$ROOT/tests/golden/weights"
head -c $((3*321*1025*4)) /dev/zero > "$OUT/plans/zero.bin"
for n in resnet18 resnet18_2D; do
  (cd "$ROOT" && python -c "from oracle import io; io.write_fp16_weights('$W/${n}_fp32.bin', '$OUT/plans/${n}_fp16.bin')")
  LD_LIBRARY_PATH="$ROOT/redtail_b200/lib" "$OUT/nvstereo_net_driver" $n 1025 321 "$W/${n}_fp32.bin" "$OUT/plans/zero.bin" "$OUT/plans/zero.bin" "$OUT/plans/${n}_1025x321_fp32.plan" dump > /dev/null
  LD_LIBRARY_PATH="$ROOT/redtail_b200/lib" "$OUT/nvstereo_net_driver" $n 1025 321 "$OUT/plans/${n}_fp16.bin" "$OUT/plans/zero.bin" "$OUT/plans/zero.bin" "$OUT/plans/${n}_1025x321_fp16.plan" dump fp16 > /dev/null
  rm -f "$OUT/plans/${n}_fp16.bin"
done
rm -f "$OUT/plans/zero.bin"
ls -la "$OUT/plans"
# The reference's sample application itself, UNCHANGED (sample_app/main.cpp: PNG in, engine, binary + 16-bit PNG out),
# against the cv:: stand-in of tools/dropin/include/opencv2 (OpenCV's C++ headers are not in this image).
$CXX $INC -I"$REF/sample_app" -c "$REF/sample_app/main.cpp" -o "$OUT/main.o"
$CXX -o "$OUT/nvstereo_sample_app" "$OUT/main.o" "$OUT"/*_net.o $LIB -lz
# The reference's TrailNet runtime, UNCHANGED (ros/packages/caffe_ros/src/tensor_net.cpp + int8_calibrator.cpp: Caffe parser ->
# buildCudaEngine -> serialize -> deserializeCudaEngine -> execute), against NvInfer.h / NvCaffeParser.h and the ros / boost / cv
# stand-ins of tools/dropin/include (neither ROS nor Boost nor OpenCV's C++ headers are in this image); driven like caffe_ros.cpp does.
CR=/root/reference/ros/packages/caffe_ros
if [ -d "$CR" ]; then
  CXX17="${CXX:-g++} -std=c++17 -O1 -w"
  $CXX17 $INC -I"$CR/include" -c "$CR/src/tensor_net.cpp" -o "$OUT/caffe_ros_tensornet.o"          # (not *_net.o: that glob is the stereo builders)
  $CXX17 $INC -I"$CR/include" -c "$CR/src/int8_calibrator.cpp" -o "$OUT/caffe_ros_int8calib.o"
  $CXX17 $INC -I"$CR/include" -c "$ROOT/tools/dropin/trailnet_driver.cpp" -o "$OUT/caffe_ros_driver.o"
  $CXX17 -o "$OUT/caffe_ros_trailnet" "$OUT/caffe_ros_driver.o" "$OUT/caffe_ros_tensornet.o" "$OUT/caffe_ros_int8calib.o" $LIB -lz
fi
echo "built: $OUT/nvstereo_tests $OUT/nvstereo_net_driver $OUT/nvstereo_sample_app $OUT/caffe_ros_trailnet"
