// nvcaffeparser1-compatible subset: what ros/packages/caffe_ros/src/tensor_net.cpp:79-124 calls to load the TrailNet model
// (`parser->parse(prototxt, caffemodel, *network, dtype)`, `blob_finder->find(name)`, `parser->destroy()`).
// Implemented in libnvstereo_inference.so (redtail_b200/csrc/host/caffe_parser.cpp): a prototxt (protobuf text) reader, a
// caffemodel (protobuf wire format, caffe.proto NetParameter.layer / BlobProto) reader and the mapping of Caffe layers onto
// INetworkDefinition::add* calls: Scale, Convolution, ReLU, Pooling (Caffe's ceil-mode output size), Eltwise SUM,
// InnerProduct, Softmax, Concat.  No protobuf library is involved.
// (The reference's sample_app/main.cpp:5 includes this header but uses nothing from it.)
#pragma once
#include "NvInfer.h"

namespace nvcaffeparser1 {

class IBlobNameToTensor {
public:
    virtual nvinfer1::ITensor* find(const char* name) const = 0;
protected:
    virtual ~IBlobNameToTensor() {}
};

class ICaffeParser {
public:
    // deploy: path of the deploy prototxt; model: path of the .caffemodel.  The parser owns the weights: keep it alive until
    // buildCudaEngine() has returned (tensor_net.cpp:161-179 destroys it after the build).  nullptr on error (logged).
    virtual const IBlobNameToTensor* parse(const char* deploy, const char* model, nvinfer1::INetworkDefinition& network,
                                           nvinfer1::DataType weightType) = 0;
    virtual void setProtobufBufferSize(size_t size) = 0;     // accepted, unused (no protobuf library)
    virtual void destroy() = 0;
protected:
    virtual ~ICaffeParser() {}
};

}  // namespace nvcaffeparser1

extern "C" void* createNvCaffeParser_INTERNAL();

namespace nvcaffeparser1 {
inline ICaffeParser* createCaffeParser() { return static_cast<ICaffeParser*>(createNvCaffeParser_INTERNAL()); }
inline void shutdownProtobufLibrary() {}
}  // namespace nvcaffeparser1
