/* Small host utilities shared by the plugin classes, the engine and the tests.  Counterpart of the reference's
 * stereoDNN/lib/internal_utils.h:28-76 (DimsUtils, StrUtils, reportError); the reference's tests include this
 * header for DimsUtils::getTensorSize and CHECKL (tests/tests_main.cpp:15,137,194).  No cuDNN here. */
#ifndef REDTAIL_INTERNAL_UTILS_H
#define REDTAIL_INTERNAL_UTILS_H

#include <NvInfer.h>
#include <cuda_runtime_api.h>

#include <cassert>
#include <string>

#include "internal_macros.h"
#include "redtail_tensorrt_plugins.h"

namespace redtail { namespace tensorrt {

using namespace nvinfer1;

class DimsUtils
{
public:
    static size_t      getTensorSize(Dims dims);          // product of d[0..nbDims)
    static Dims        getStrides(Dims dims);             // dense row-major strides, in elements
    static bool        areEqual(Dims d1, Dims d2);
    static std::string toString(Dims dims);               // "{a, b, c}"
    DimsUtils(DimsUtils&&) = delete;
};

class StrUtils
{
public:
    static std::string toString(DataType type);
    static std::string toString(PluginFormat format);
    StrUtils(StrUtils&&) = delete;
};

/* Logs "file:line: func: CUDA error N (name: text)." at kERROR and asserts (reference: lib/internal_utils.cpp:260-270).
 * `int` overload: status codes of the redtail_b200 C-ABI (cudaError_t values or negative rt_status). */
void reportError(cudaError_t status, const char* file, int line, const char* func, ILogger& log);
void reportError(int status, const char* file, int line, const char* func, ILogger& log);

} }

#endif
