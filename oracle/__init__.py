"""CPU oracle for the redtail stereoDNN plugin path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy / PyTorch-CPU), the arithmetic of the reference's
TensorRT plugin library `stereoDNN/lib` and of the TRT-native layers its generated network
builders use.  It is the parity checker for the CUDA path in `redtail_b200/`.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import it; nothing under `redtail_b200/` does;
  * it never touches CUDA.

Pinning: every op below reproduces the reference's own 58 known-answer tensors
(`stereoDNN/tests/data/*.bin`, re-packed in `tests/golden/plugin_fixtures.npz`) within the
tolerances the reference's gtest suite uses (`stereoDNN/tests/tests_main.cpp:280-1099`);
see `tests/test_oracle_golden.py`.  Network-level disparity has no fixture in the reference
(no golden disparity map ships); it is pinned transitively: fixture-pinned ops + the
reference's trained weights + the reference's wiring (`sample_app/*_net.cpp`).

The reference implementation itself cannot be built here: it needs the TensorRT 3/4
`IPlugin` API, cuDNN 7 and a GPU ("Oracle").
"""
