// boost::filesystem stand-in over std::filesystem for ros/packages/caffe_ros/src/int8_calibrator.cpp (Boost is not in this image).
#pragma once
#include <filesystem>
#include <fstream>      // boost/filesystem.hpp pulls <fstream> in (boost/filesystem/fstream.hpp); int8_calibrator.cpp:91,108 relies on that
namespace boost {
namespace filesystem {
using std::filesystem::directory_iterator;
using std::filesystem::exists;
using std::filesystem::is_directory;
using std::filesystem::is_regular_file;
using std::filesystem::path;
}  // namespace filesystem
}  // namespace boost
