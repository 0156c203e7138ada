// nvcaffeparser1::ICaffeParser for the layer types of the TrailNet S-ResNet-18 classifier -- what
// ros/packages/caffe_ros/src/tensor_net.cpp:79-124 gets from TensorRT's Caffe parser (`parser->parse(prototxt, caffemodel,
// *network, dtype)`, then `blob_finder->find(output_blob)`).  Two small readers, no protobuf library:
//   * the deploy prototxt: protobuf TEXT format -> a tree of (name, scalar | message) fields;
//   * the .caffemodel: protobuf WIRE format, caffe.proto NetParameter.layer (field 100) -> name (1), blobs (7);
//     BlobProto: shape (7){dim (1)}, data (5, packed floats), legacy num/channels/height/width (1-4).
// Caffe layer -> INetworkDefinition call (Caffe semantics, BVLC caffe 1.0 layer definitions):
//   Scale        addScale(kCHANNEL, shift = bias blob, scale = scale blob)        (blobs from the model, else the fillers)
//   Convolution  addConvolution + setStride + setPadding
//   ReLU         addActivation(kRELU)
//   Pooling      addPooling(kMAX | kAVERAGE) with Caffe's output extent: ceil((in + 2 pad - k) / stride) + 1, minus one if the
//                last window would start outside the padded image (pooling_layer.cpp) -- installed as the network's pooling formula
//   Eltwise      addElementWise(kSUM)
//   InnerProduct addFullyConnected
//   Softmax      addSoftMax
//   Concat       addConcatenation (axis 1)
// In-place layers (top == bottom) simply rebind the blob name to the new tensor, as the Caffe parser does.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "NvCaffeParser.h"

using namespace nvinfer1;

namespace {

// ---- protobuf text format ------------------------------------------------------------------------------------------
struct Msg;
struct Field {
    std::string name;
    std::string scalar;               // valid when msg == nullptr
    std::shared_ptr<Msg> msg;
};
struct Msg {
    std::vector<Field> fields;
    const Field* first(const std::string& n) const
    {
        for (const auto& f : fields) if (f.name == n) return &f;
        return nullptr;
    }
    std::vector<const Field*> all(const std::string& n) const
    {
        std::vector<const Field*> v;
        for (const auto& f : fields) if (f.name == n) v.push_back(&f);
        return v;
    }
    std::string str(const std::string& n, const std::string& dflt = "") const
    {
        const Field* f = first(n);
        return f && !f->msg ? f->scalar : dflt;
    }
    double num(const std::string& n, double dflt) const
    {
        const Field* f = first(n);
        return f && !f->msg ? atof(f->scalar.c_str()) : dflt;
    }
    const Msg* sub(const std::string& n) const
    {
        const Field* f = first(n);
        return f && f->msg ? f->msg.get() : nullptr;
    }
};

struct TextParser {
    const std::string& s;
    size_t pos = 0;
    std::string err;
    explicit TextParser(const std::string& text) : s(text) {}
    void skip()
    {
        while (pos < s.size()) {
            if (isspace(static_cast<unsigned char>(s[pos]))) ++pos;
            else if (s[pos] == '#') { while (pos < s.size() && s[pos] != '\n') ++pos; }
            else break;
        }
    }
    std::string token()
    {
        skip();
        if (pos >= s.size()) return std::string();
        const char c = s[pos];
        if (c == '{' || c == '}' || c == ':') { ++pos; return std::string(1, c); }
        if (c == '"' || c == '\'') {
            std::string out(1, '"');
            ++pos;
            while (pos < s.size() && s[pos] != c) { if (s[pos] == '\\' && pos + 1 < s.size()) ++pos; out.push_back(s[pos++]); }
            ++pos;
            return out;                    // leading '"' marks a string literal
        }
        const size_t b = pos;
        while (pos < s.size() && !isspace(static_cast<unsigned char>(s[pos])) && s[pos] != '{' && s[pos] != '}' && s[pos] != ':' && s[pos] != '#') ++pos;
        return s.substr(b, pos - b);
    }
    bool message(Msg& m, bool top)
    {
        for (;;) {
            const std::string name = token();
            if (name.empty()) { if (!top) err = "prototxt: missing '}'"; return top; }
            if (name == "}") { if (top) { err = "prototxt: unbalanced '}'"; return false; } return true; }
            Field f;
            f.name = name;
            std::string t = token();
            if (t == ":") t = token();
            if (t == "{") {
                f.msg.reset(new Msg());
                if (!message(*f.msg, false)) return false;
            } else if (t.empty() || t == "}" || t == ":") {
                err = "prototxt: value expected after '" + name + "'";
                return false;
            } else f.scalar = t[0] == '"' ? t.substr(1) : t;
            m.fields.push_back(std::move(f));
        }
    }
};

// ---- protobuf wire format --------------------------------------------------------------------------------------------
struct Blob { std::vector<int64_t> dims; std::vector<float> data; };

struct WireReader {
    const uint8_t* p; size_t n; size_t pos = 0; bool ok = true;
    WireReader(const uint8_t* b, size_t len) : p(b), n(len) {}
    uint64_t varint()
    {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (pos >= n) { ok = false; return 0; }
            const uint8_t b = p[pos++];
            v |= static_cast<uint64_t>(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    // next field: number, wire type, and for length-delimited fields its span.  false at the end / on error.
    bool next(int& fno, int& wt, uint64_t& val, const uint8_t*& sb, size_t& sl)
    {
        if (pos >= n || !ok) return false;
        const uint64_t key = varint();
        if (!ok) return false;
        fno = static_cast<int>(key >> 3); wt = static_cast<int>(key & 7);
        val = 0; sb = nullptr; sl = 0;
        switch (wt) {
            case 0: val = varint(); break;
            case 1: if (pos + 8 > n) { ok = false; return false; } sb = p + pos; sl = 8; pos += 8; break;
            case 2: { const uint64_t l = varint(); if (!ok || pos + l > n) { ok = false; return false; } sb = p + pos; sl = l; pos += l; break; }
            case 5: if (pos + 4 > n) { ok = false; return false; } sb = p + pos; sl = 4; pos += 4; break;
            default: ok = false; return false;
        }
        return ok;
    }
};

bool parseBlob(const uint8_t* b, size_t len, Blob& out)
{
    WireReader r(b, len);
    int fno, wt; uint64_t val; const uint8_t* sb; size_t sl;
    int64_t legacy[5] = {0, 0, 0, 0, 0};
    bool has_legacy = false, has_shape = false;
    while (r.next(fno, wt, val, sb, sl)) {
        if (fno == 7 && wt == 2) {                                    // BlobShape
            has_shape = true;
            WireReader r2(sb, sl);
            int f2, w2; uint64_t v2; const uint8_t* s2; size_t l2;
            while (r2.next(f2, w2, v2, s2, l2)) {
                if (f2 == 1 && w2 == 2) { WireReader r3(s2, l2); while (r3.pos < r3.n && r3.ok) out.dims.push_back(static_cast<int64_t>(r3.varint())); }
                else if (f2 == 1 && w2 == 0) out.dims.push_back(static_cast<int64_t>(v2));
            }
            if (!r2.ok) return false;
        } else if (fno == 5 && wt == 2) {                             // packed float data
            out.data.resize(sl / 4);
            memcpy(out.data.data(), sb, out.data.size() * 4);
        } else if (fno == 5 && wt == 5) {
            float f;
            memcpy(&f, sb, 4);
            out.data.push_back(f);
        } else if (fno >= 1 && fno <= 4 && wt == 0) { legacy[fno] = static_cast<int64_t>(val); has_legacy = true; }
    }
    if (!r.ok) return false;
    if (!has_shape) {
        if (has_legacy) for (int i = 1; i <= 4; ++i) out.dims.push_back(legacy[i] > 0 ? legacy[i] : 1);
        else out.dims.push_back(static_cast<int64_t>(out.data.size()));
    }
    int64_t vol = 1;
    for (int64_t d : out.dims) vol *= d;
    return vol == static_cast<int64_t>(out.data.size());
}

bool parseCaffeModel(const std::string& bytes, std::map<std::string, std::vector<Blob>>& out)
{
    WireReader r(reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size());
    int fno, wt; uint64_t val; const uint8_t* sb; size_t sl;
    while (r.next(fno, wt, val, sb, sl)) {
        if (fno != 100 || wt != 2) continue;                          // NetParameter.layer (LayerParameter)
        WireReader r2(sb, sl);
        int f2, w2; uint64_t v2; const uint8_t* s2; size_t l2;
        std::string name;
        std::vector<Blob> blobs;
        while (r2.next(f2, w2, v2, s2, l2)) {
            if (f2 == 1 && w2 == 2) name.assign(reinterpret_cast<const char*>(s2), l2);
            else if (f2 == 7 && w2 == 2) {
                Blob b;
                if (!parseBlob(s2, l2, b)) return false;
                blobs.push_back(std::move(b));
            }
        }
        if (!r2.ok) return false;
        if (!name.empty() && !blobs.empty()) out[name] = std::move(blobs);
    }
    return r.ok;
}

bool readFile(const char* path, std::string& out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

// Caffe's pooling output extent (pooling_layer.cpp:Reshape).
struct CaffePoolingFormula : public IOutputDimensionsFormula {
    DimsHW compute(DimsHW in, DimsHW k, DimsHW stride, DimsHW pad, DimsHW, const char*) const override
    {
        auto one = [](int i, int kk, int s, int p) {
            int o = static_cast<int>(std::ceil(static_cast<double>(i + 2 * p - kk) / s)) + 1;
            if (p > 0 && (o - 1) * s >= i + p) --o;
            return o;
        };
        return DimsHW(one(in.h(), k.h(), stride.h(), pad.h()), one(in.w(), k.w(), stride.w(), pad.w()));
    }
};

class BlobMap : public nvcaffeparser1::IBlobNameToTensor {
public:
    ITensor* find(const char* name) const override
    {
        auto it = map.find(name ? name : "");
        return it == map.end() ? nullptr : it->second;
    }
    std::map<std::string, ITensor*> map;
};

class CaffeParserImpl : public nvcaffeparser1::ICaffeParser {
public:
    const nvcaffeparser1::IBlobNameToTensor* parse(const char* deploy, const char* model, INetworkDefinition& network, DataType weightType) override
    {
        (void)weightType;       // weights stay fp32 on the host; the engine picks the arithmetic (builder->setHalf2Mode)
        std::string text, bin;
        if (!deploy || !readFile(deploy, text)) return error(std::string("cannot read prototxt ") + (deploy ? deploy : "(null)"));
        if (!model || !readFile(model, bin)) return error(std::string("cannot read caffemodel ") + (model ? model : "(null)"));
        Msg net;
        TextParser tp(text);
        if (!tp.message(net, true)) return error(tp.err);
        if (!parseCaffeModel(bin, blobs_)) return error("malformed caffemodel (protobuf wire format)");
        network.setPoolingOutputDimensionsFormula(&pool_formula_);

        // inputs: `input: "data"` + input_shape { dim x4 } (or input_dim x4)
        const std::string in_name = net.str("input", "data");
        std::vector<int> idims;
        if (const Msg* sh = net.sub("input_shape")) for (const Field* f : sh->all("dim")) idims.push_back(atoi(f->scalar.c_str()));
        else for (const Field* f : net.all("input_dim")) idims.push_back(atoi(f->scalar.c_str()));
        if (idims.size() != 4) return error("prototxt: a 4-D input_shape is required");
        ITensor* in = network.addInput(in_name.c_str(), DataType::kFLOAT, DimsCHW(idims[1], idims[2], idims[3]));
        if (!in) return error("addInput failed");
        map_.map[in_name] = in;

        for (const Field* lf : net.all("layer")) {
            if (!lf->msg) continue;
            const Msg& l = *lf->msg;
            const std::string name = l.str("name"), type = l.str("type");
            std::vector<ITensor*> bottoms;
            for (const Field* b : l.all("bottom")) {
                ITensor* t = map_.find(b->scalar.c_str());
                if (!t) return error(name + ": unknown bottom blob '" + b->scalar + "'");
                bottoms.push_back(t);
            }
            const std::string top = l.str("top");
            if (bottoms.empty() || top.empty()) return error(name + ": layers need a bottom and a top");
            const Dims bd = bottoms[0]->getDimensions();
            ILayer* layer = nullptr;
            if (type == "Scale") {
                const Msg* sp = l.sub("scale_param");
                const int c = bd.d[0];
                const bool bias_term = sp && sp->str("bias_term", "false") == "true";
                Weights scale = blobOrFill(name, 0, c, sp && sp->sub("filler") ? static_cast<float>(sp->sub("filler")->num("value", 1.0)) : 1.f);
                Weights shift{DataType::kFLOAT, nullptr, 0};
                if (bias_term) shift = blobOrFill(name, 1, c, sp->sub("bias_filler") ? static_cast<float>(sp->sub("bias_filler")->num("value", 0.0)) : 0.f);
                if (scale.count != c || (bias_term && shift.count != c)) return error(name + ": scale blob size != channels");
                layer = network.addScale(*bottoms[0], ScaleMode::kCHANNEL, shift, scale, Weights{DataType::kFLOAT, nullptr, 0});
            } else if (type == "Convolution") {
                const Msg* cp = l.sub("convolution_param");
                if (!cp) return error(name + ": convolution_param missing");
                const int k = static_cast<int>(cp->num("kernel_size", 1)), st = static_cast<int>(cp->num("stride", 1)), pd = static_cast<int>(cp->num("pad", 0));
                const int maps = static_cast<int>(cp->num("num_output", 0));
                if (cp->num("group", 1) != 1 || cp->num("dilation", 1) != 1) return error(name + ": grouped / dilated convolutions are not supported");
                auto it = blobs_.find(name);
                if (it == blobs_.end() || it->second.empty()) return error(name + ": no weights in the caffemodel");
                const Blob& w = it->second[0];
                if (static_cast<int64_t>(w.data.size()) != static_cast<int64_t>(maps) * bd.d[0] * k * k) return error(name + ": weight blob does not match the layer shape");
                Weights kw{DataType::kFLOAT, w.data.data(), static_cast<int64_t>(w.data.size())};
                Weights bw{DataType::kFLOAT, nullptr, 0};
                if (cp->str("bias_term", "true") == "true" && it->second.size() > 1)
                    bw = Weights{DataType::kFLOAT, it->second[1].data.data(), static_cast<int64_t>(it->second[1].data.size())};
                auto* cl = network.addConvolution(*bottoms[0], maps, DimsHW(k, k), kw, bw);
                if (cl) { cl->setStride(DimsHW(st, st)); cl->setPadding(DimsHW(pd, pd)); }
                layer = cl;
            } else if (type == "ReLU") {
                if (l.sub("relu_param") && l.sub("relu_param")->num("negative_slope", 0) != 0) return error(name + ": leaky ReLU is not supported");
                layer = network.addActivation(*bottoms[0], ActivationType::kRELU);
            } else if (type == "Pooling") {
                const Msg* pp = l.sub("pooling_param");
                if (!pp) return error(name + ": pooling_param missing");
                const std::string pool = pp->str("pool", "MAX");
                if (pool != "MAX" && pool != "AVE") return error(name + ": pooling method " + pool + " is not supported");
                const int k = static_cast<int>(pp->num("kernel_size", 1)), st = static_cast<int>(pp->num("stride", 1)), pd = static_cast<int>(pp->num("pad", 0));
                auto* pl = network.addPooling(*bottoms[0], pool == "MAX" ? PoolingType::kMAX : PoolingType::kAVERAGE, DimsHW(k, k));
                if (pl) { pl->setStride(DimsHW(st, st)); pl->setPadding(DimsHW(pd, pd)); }
                layer = pl;
            } else if (type == "Eltwise") {
                if (bottoms.size() != 2) return error(name + ": Eltwise with two bottoms only");
                if (l.sub("eltwise_param") && l.sub("eltwise_param")->str("operation", "SUM") != "SUM") return error(name + ": Eltwise SUM only");
                layer = network.addElementWise(*bottoms[0], *bottoms[1], ElementWiseOperation::kSUM);
            } else if (type == "InnerProduct") {
                const Msg* ip = l.sub("inner_product_param");
                const int maps = ip ? static_cast<int>(ip->num("num_output", 0)) : 0;
                auto it = blobs_.find(name);
                if (maps <= 0 || it == blobs_.end() || it->second.empty()) return error(name + ": no weights in the caffemodel");
                const Blob& w = it->second[0];
                Weights kw{DataType::kFLOAT, w.data.data(), static_cast<int64_t>(w.data.size())};
                Weights bw{DataType::kFLOAT, nullptr, 0};
                if (it->second.size() > 1) bw = Weights{DataType::kFLOAT, it->second[1].data.data(), static_cast<int64_t>(it->second[1].data.size())};
                layer = network.addFullyConnected(*bottoms[0], maps, kw, bw);
            } else if (type == "Softmax") {
                layer = network.addSoftMax(*bottoms[0]);
            } else if (type == "Concat") {
                if (l.sub("concat_param") && l.sub("concat_param")->num("axis", 1) != 1) return error(name + ": Concat along the channel axis only");
                layer = network.addConcatenation(bottoms.data(), static_cast<int>(bottoms.size()));
            } else {
                return error(name + ": Caffe layer type '" + type + "' is not supported");
            }
            if (!layer || !layer->getOutput(0)) return error(name + ": the network rejected the layer");
            layer->setName(name.c_str());
            layer->getOutput(0)->setName(top.c_str());
            if (layer->getOutput(0)->getDimensions().nbDims == 0) return error(name + ": could not resolve the output shape");
            map_.map[top] = layer->getOutput(0);
        }
        return &map_;
    }
    void setProtobufBufferSize(size_t) override {}
    void destroy() override { delete this; }

private:
    const nvcaffeparser1::IBlobNameToTensor* error(const std::string& what)
    {
        fprintf(stderr, "[redtail caffe parser] %s\n", what.c_str());
        return nullptr;
    }
    // The layer's idx-th blob from the model, or `c` copies of the prototxt filler value.
    Weights blobOrFill(const std::string& layer, size_t idx, int c, float fill)
    {
        auto it = blobs_.find(layer);
        if (it != blobs_.end() && it->second.size() > idx)
            return Weights{DataType::kFLOAT, it->second[idx].data.data(), static_cast<int64_t>(it->second[idx].data.size())};
        fills_.emplace_back(static_cast<size_t>(c), fill);
        return Weights{DataType::kFLOAT, fills_.back().data(), c};
    }
    std::map<std::string, std::vector<Blob>> blobs_;      // owns every weight the network points to
    std::vector<std::vector<float>> fills_;
    BlobMap map_;
    CaffePoolingFormula pool_formula_;
};

}  // namespace

extern "C" void* createNvCaffeParser_INTERNAL() { return static_cast<nvcaffeparser1::ICaffeParser*>(new CaffeParserImpl()); }
