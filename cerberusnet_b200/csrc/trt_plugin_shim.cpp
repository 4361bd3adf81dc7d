// trt_plugin_shim.cpp -- TensorRT IPluginV2DynamicExt shim over the C ABI.
//
// Same plugin type/version ("correlation", "1"), field names, 24-byte serialisation, output
// dimension rule and enqueue signature as the reference plugin
// (runtime/cerberus_net/trt_plugins/correlation.{hpp,cpp,cu}), so an engine built from the
// reference's ONNX export (utilities/onnx_export.py:18-23) binds to it unchanged.  Differences, all
// on purpose: zero workspace (reference: 2*N*C*(H+2p)*(W+2p)*sizeof(T), then over-run,
// correlation.cu:112,137), no private streams and no host synchronisation inside enqueue
// (reference: three cudaStreamSynchronize, correlation.cu:105,124-125), so the call is
// CUDA-graph capturable; errors come back as the return value instead of abort()/throw.
//
// TensorRT is not part of this repository's image: the file compiles to nothing unless
// <NvInfer.h> is on the include path (add it to the reference's trt_plugins_lib next to
// libcerberus_costvolume.so).  The adapter itself (cerb_trt_corr_enqueue) is tested through a
// layout-compatible mock of PluginTensorDesc in tests/test_gpu_parity.py.
#if defined(__has_include)
#if __has_include(<NvInfer.h>)
#define CERB_HAVE_TENSORRT 1
#endif
#endif

#ifdef CERB_HAVE_TENSORRT
#include <NvInfer.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/cerberus_trt_plugin.h"

namespace cerb_trt {

static_assert(sizeof(nvinfer1::PluginTensorDesc) == sizeof(cerb_trt_tensor_desc) ||
                  sizeof(nvinfer1::PluginTensorDesc) == sizeof(cerb_trt_tensor_desc64),
              "PluginTensorDesc layout changed: adapt cerb_trt_tensor_desc");

class CorrelationPlugin : public nvinfer1::IPluginV2DynamicExt {
 public:
  explicit CorrelationPlugin(const nvinfer1::PluginFieldCollection& fc) {
    cerb_trt_corr_default_fields(&f_);  // correlation.cpp:54-62
    for (int i = 0; i < fc.nbFields; ++i) {
      const char* n = fc.fields[i].name;
      const int v = *static_cast<const int*>(fc.fields[i].data);
      if (!strcmp(n, "pad_size")) f_.pad_size = v;
      else if (!strcmp(n, "kernel_size")) f_.kernel_size = v;
      else if (!strcmp(n, "max_displacement")) f_.max_displacement = v;
      else if (!strcmp(n, "stride1")) f_.stride1 = v;
      else if (!strcmp(n, "stride2")) f_.stride2 = v;
      else if (!strcmp(n, "corr_multiply")) f_.corr_multiply = v;
    }
  }
  CorrelationPlugin(const void* data, size_t length) { cerb_trt_corr_deserialize(data, length, &f_); }

  int getNbOutputs() const noexcept override { return 1; }
  nvinfer1::DimsExprs getOutputDimensions(int, const nvinfer1::DimsExprs* in, int, nvinfer1::IExprBuilder& eb) noexcept override {
    const int kr = (f_.kernel_size - 1) / 2, border = kr + f_.max_displacement;
    const int d = (f_.max_displacement / f_.stride2) * 2 + 1;
    nvinfer1::DimsExprs o;
    o.nbDims = 4;
    o.d[0] = in[0].d[0];
    o.d[1] = eb.constant(d * d);
    for (int k = 2; k < 4; ++k)  // ceil((H + 2p - 2*border) / s1), correlation.cpp:178-205
      o.d[k] = eb.operation(nvinfer1::DimensionOperation::kCEIL_DIV,
                            *eb.operation(nvinfer1::DimensionOperation::kSUB,
                                          *eb.operation(nvinfer1::DimensionOperation::kSUM, *in[0].d[k], *eb.constant(2 * f_.pad_size)),
                                          *eb.constant(2 * border)),
                            *eb.constant(f_.stride1));
    return o;
  }
  int initialize() noexcept override { return 0; }   // no private streams
  void terminate() noexcept override {}
  size_t getWorkspaceSize(const nvinfer1::PluginTensorDesc*, int, const nvinfer1::PluginTensorDesc*, int) const noexcept override { return 0; }
  int enqueue(const nvinfer1::PluginTensorDesc* inputDesc, const nvinfer1::PluginTensorDesc* outputDesc,
              const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept override {
    if (sizeof(nvinfer1::PluginTensorDesc) == sizeof(cerb_trt_tensor_desc))
      return cerb_trt_corr_enqueue(&f_, reinterpret_cast<const cerb_trt_tensor_desc*>(inputDesc),
                                   reinterpret_cast<const cerb_trt_tensor_desc*>(outputDesc), inputs, outputs, workspace, stream);
    return cerb_trt_corr_enqueue_i64(&f_, reinterpret_cast<const cerb_trt_tensor_desc64*>(inputDesc),
                                     reinterpret_cast<const cerb_trt_tensor_desc64*>(outputDesc), inputs, outputs, workspace, stream);
  }
  void configurePlugin(const nvinfer1::DynamicPluginTensorDesc*, int, const nvinfer1::DynamicPluginTensorDesc*, int) noexcept override {}
  size_t getSerializationSize() const noexcept override { return cerb_trt_corr_serialization_size(); }
  void serialize(void* buffer) const noexcept override { cerb_trt_corr_serialize(&f_, buffer); }
  bool supportsFormatCombination(int pos, const nvinfer1::PluginTensorDesc* io, int nbIn, int nbOut) noexcept override {
    if (sizeof(nvinfer1::PluginTensorDesc) != sizeof(cerb_trt_tensor_desc))
      return io[pos].format == nvinfer1::TensorFormat::kLINEAR &&
             (io[pos].type == nvinfer1::DataType::kFLOAT || io[pos].type == nvinfer1::DataType::kHALF) && io[pos].type == io[0].type;
    return cerb_trt_corr_supports_format(pos, reinterpret_cast<const cerb_trt_tensor_desc*>(io), nbIn, nbOut) != 0;
  }
  const char* getPluginType() const noexcept override { return CERB_TRT_CORR_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_CORR_PLUGIN_VERSION; }
  void destroy() noexcept override { delete this; }
  nvinfer1::IPluginV2DynamicExt* clone() const noexcept override {
    auto* p = new CorrelationPlugin(*this);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }
  nvinfer1::DataType getOutputDataType(int, const nvinfer1::DataType* t, int) const noexcept override { return t[0]; }

 private:
  cerb_trt_corr_fields f_{};
  std::string ns_;
};

class CorrelationPluginCreator : public nvinfer1::IPluginCreator {
 public:
  CorrelationPluginCreator() {
    static const char* names[6] = {"pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply"};
    attrs_.clear();
    for (const char* n : names) attrs_.emplace_back(nvinfer1::PluginField(n, nullptr, nvinfer1::PluginFieldType::kINT32, 1));
    fc_.nbFields = (int)attrs_.size();
    fc_.fields = attrs_.data();
  }
  const char* getPluginName() const noexcept override { return CERB_TRT_CORR_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_CORR_PLUGIN_VERSION; }
  const nvinfer1::PluginFieldCollection* getFieldNames() noexcept override { return &fc_; }
  nvinfer1::IPluginV2* createPlugin(const char*, const nvinfer1::PluginFieldCollection* fc) noexcept override {
    auto* p = new CorrelationPlugin(*fc);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  nvinfer1::IPluginV2* deserializePlugin(const char*, const void* data, size_t len) noexcept override {
    auto* p = new CorrelationPlugin(data, len);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }

 private:
  nvinfer1::PluginFieldCollection fc_{};
  std::vector<nvinfer1::PluginField> attrs_;
  std::string ns_;
};

REGISTER_TENSORRT_PLUGIN(CorrelationPluginCreator);

}  // namespace cerb_trt
#endif  // CERB_HAVE_TENSORRT
