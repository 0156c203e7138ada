// boost::iequals stand-in (ros/packages/caffe_ros/src/tensor_net.cpp:338-352); Boost is not in this image.
#pragma once
#include <cctype>
#include <string>
namespace boost {
inline bool iequals(const std::string& a, const std::string& b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (std::tolower(static_cast<unsigned char>(a[i])) != std::tolower(static_cast<unsigned char>(b[i]))) return false;
    return true;
}
}  // namespace boost
