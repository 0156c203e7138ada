// DimsUtils / StrUtils / reportError (see include/internal_utils.h).
#include "internal_utils.h"

#include <sstream>

namespace redtail { namespace tensorrt {

size_t DimsUtils::getTensorSize(Dims dims)
{
    size_t n = 1;
    for (int i = 0; i < dims.nbDims; i++) n *= static_cast<size_t>(dims.d[i]);
    return n;
}

Dims DimsUtils::getStrides(Dims dims)
{
    Dims s = dims;
    int64_t acc = 1;
    for (int i = dims.nbDims - 1; i >= 0; i--) {
        s.d[i] = static_cast<int>(acc);
        acc *= dims.d[i];
    }
    return s;
}

bool DimsUtils::areEqual(Dims d1, Dims d2)
{
    if (d1.nbDims != d2.nbDims) return false;
    for (int i = 0; i < d1.nbDims; i++)
        if (d1.d[i] != d2.d[i]) return false;
    return true;
}

std::string DimsUtils::toString(Dims dims)
{
    std::ostringstream s;
    s << "{";
    for (int i = 0; i < dims.nbDims; i++) s << (i ? ", " : "") << dims.d[i];
    s << "}";
    return s.str();
}

std::string StrUtils::toString(DataType type)
{
    switch (type) {
        case DataType::kFLOAT: return "FLOAT";
        case DataType::kHALF:  return "HALF";
        case DataType::kINT8:  return "INT8";
        case DataType::kINT32: return "INT32";
    }
    return "UNKNOWN";
}

std::string StrUtils::toString(PluginFormat format)
{
    switch (format) {
        case PluginFormat::kNCHW:   return "NCHW";
        case PluginFormat::kNC2HW2: return "NC2HW2";
        case PluginFormat::kNHWC8:  return "NHWC8";
    }
    return "UNKNOWN";
}

void reportError(cudaError_t status, const char* file, int line, const char* func, ILogger& log)
{
    if (status == cudaSuccess) return;
    std::ostringstream s;
    s << file << ":" << line << ": " << func << ": CUDA error " << static_cast<int>(status) << " ("
      << cudaGetErrorName(status) << ": " << cudaGetErrorString(status) << ").";
    log.log(ILogger::Severity::kERROR, s.str().c_str());
    assert(status == cudaSuccess);
}

void reportError(int status, const char* file, int line, const char* func, ILogger& log)
{
    if (status == 0) return;
    if (status > 0) {
        reportError(static_cast<cudaError_t>(status), file, line, func, log);
        return;
    }
    std::ostringstream s;
    s << file << ":" << line << ": " << func << ": redtail_b200 error " << status
      << (status == -1 ? " (bad argument)." : status == -2 ? " (unsupported configuration)." : ".");
    log.log(ILogger::Severity::kERROR, s.str().c_str());
    assert(status == 0);
}

} }
