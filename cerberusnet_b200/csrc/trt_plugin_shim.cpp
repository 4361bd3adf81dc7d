// trt_plugin_shim.cpp -- TensorRT IPluginV2DynamicExt plugins over the C ABI.
//
//   CorrelationPlugin      type "correlation"      v"1"  -- same type/version, field names, 24-byte serialisation, output
//                          dimension rule and enqueue signature as the reference plugin
//                          (runtime/cerberus_net/trt_plugins/correlation.{hpp,cpp,cu}), so an engine built from the
//                          reference's ONNX export (utilities/onnx_export.py:18-23) binds to it unchanged.
//   GridSamplerPlugin      type "grid_sampler"     v"1"  -- trt_plugins/grid_sampler.{hpp,cpp,cu}: fields align_corners /
//                          interpolation_mode / padding_mode, 9-byte serialisation, the plugin's own un-normalise.
//   WarpCorrelationPlugin  type "warp_correlation" v"1"  -- NEW fused node {im1, im2, flow} -> activated cost volume; replaces
//                          per pyramid level 2x ScatterND + Transpose + grid_sampler + correlation + LeakyRelu
//                          (SURVEY.md 3.3); exported by cerberusnet_b200.onnx_export.
//
// Differences from the reference plugins, all on purpose: zero workspace (reference: 2*N*C*(H+2p)*(W+2p)*sizeof(T),
// then over-run, correlation.cu:112,137), no private streams and no host synchronisation inside enqueue (reference:
// three cudaStreamSynchronize, correlation.cu:105,124-125), so the call is CUDA-graph capturable; errors come back
// as the return value instead of abort()/throw.
//
// TensorRT is not part of this repository's image: the file compiles to nothing unless <NvInfer.h> is on the
// include path (add it to the reference's trt_plugins_lib next to libcerberus_costvolume.so).
// tests/test_trt_shim.py compiles it against tests/trt_stub/NvInfer.h -- a minimal stand-in for the plugin API --
// and drives every class through its creator (field parsing, output dimensions, serialisation round trip, clone,
// format support) on the CPU and through enqueue on the GPU.
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
constexpr bool kDesc32 = sizeof(nvinfer1::PluginTensorDesc) == sizeof(cerb_trt_tensor_desc);

static int field_int(const nvinfer1::PluginField& f) { return *static_cast<const int*>(f.data); }

static void parse_corr_fields(const nvinfer1::PluginFieldCollection& fc, cerb_trt_corr_fields& f) {
  cerb_trt_corr_default_fields(&f);  // correlation.cpp:54-62
  for (int i = 0; i < fc.nbFields; ++i) {
    const char* n = fc.fields[i].name;
    if (!n || !fc.fields[i].data) continue;
    if (!strcmp(n, "pad_size")) f.pad_size = field_int(fc.fields[i]);
    else if (!strcmp(n, "kernel_size")) f.kernel_size = field_int(fc.fields[i]);
    else if (!strcmp(n, "max_displacement")) f.max_displacement = field_int(fc.fields[i]);
    else if (!strcmp(n, "stride1")) f.stride1 = field_int(fc.fields[i]);
    else if (!strcmp(n, "stride2")) f.stride2 = field_int(fc.fields[i]);
    else if (!strcmp(n, "corr_multiply")) f.corr_multiply = field_int(fc.fields[i]);
  }
}

// [N, D*D, ceil((H + 2p - 2*border)/s1), ceil((W + 2p - 2*border)/s1)], correlation.cpp:178-205
static nvinfer1::DimsExprs corr_output_dims(const cerb_trt_corr_fields& f, const nvinfer1::DimsExprs* in, nvinfer1::IExprBuilder& eb) {
  const int kr = (f.kernel_size - 1) / 2, border = kr + f.max_displacement;
  const int d = (f.max_displacement / f.stride2) * 2 + 1;
  nvinfer1::DimsExprs o;
  o.nbDims = 4;
  o.d[0] = in[0].d[0];
  o.d[1] = eb.constant(d * d);
  for (int k = 2; k < 4; ++k)
    o.d[k] = eb.operation(nvinfer1::DimensionOperation::kCEIL_DIV,
                          *eb.operation(nvinfer1::DimensionOperation::kSUB,
                                        *eb.operation(nvinfer1::DimensionOperation::kSUM, *in[0].d[k], *eb.constant(2 * f.pad_size)),
                                        *eb.constant(2 * border)),
                          *eb.constant(f.stride1));
  return o;
}

// Everything the three plugins share: no state beyond the fields, no streams, no workspace.
class PluginBase : public nvinfer1::IPluginV2DynamicExt {
 public:
  int getNbOutputs() const noexcept override { return 1; }
  int initialize() noexcept override { return 0; }   // no private streams (reference: correlation.cpp:107-112)
  void terminate() noexcept override {}
  size_t getWorkspaceSize(const nvinfer1::PluginTensorDesc*, int, const nvinfer1::PluginTensorDesc*, int) const noexcept override { return 0; }
  void configurePlugin(const nvinfer1::DynamicPluginTensorDesc*, int, const nvinfer1::DynamicPluginTensorDesc*, int) noexcept override {}
  void destroy() noexcept override { delete this; }
  void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }
  nvinfer1::DataType getOutputDataType(int, const nvinfer1::DataType* t, int) const noexcept override { return t[0]; }

 protected:
  std::string ns_;
};

// linear format, kFLOAT or kHALF, same type as input 0 (flow input of the fused node: kFLOAT)
static bool linear_float_or_half(const nvinfer1::PluginTensorDesc* io, int pos, int flow_pos = -1) {
  if (io[pos].format != nvinfer1::TensorFormat::kLINEAR) return false;
  if (pos == flow_pos) return io[pos].type == nvinfer1::DataType::kFLOAT;
  return (io[pos].type == nvinfer1::DataType::kFLOAT || io[pos].type == nvinfer1::DataType::kHALF) && io[pos].type == io[0].type;
}

// ------------------------------------------------------------------ correlation ----------
class CorrelationPlugin : public PluginBase {
 public:
  explicit CorrelationPlugin(const nvinfer1::PluginFieldCollection& fc) { parse_corr_fields(fc, f_); }
  CorrelationPlugin(const void* data, size_t length) {
    if (cerb_trt_corr_deserialize(data, length, &f_) != 0) cerb_trt_corr_default_fields(&f_);
  }
  nvinfer1::DimsExprs getOutputDimensions(int, const nvinfer1::DimsExprs* in, int, nvinfer1::IExprBuilder& eb) noexcept override {
    return corr_output_dims(f_, in, eb);
  }
  int enqueue(const nvinfer1::PluginTensorDesc* inputDesc, const nvinfer1::PluginTensorDesc* outputDesc,
              const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept override {
    if (kDesc32)
      return cerb_trt_corr_enqueue(&f_, reinterpret_cast<const cerb_trt_tensor_desc*>(inputDesc),
                                   reinterpret_cast<const cerb_trt_tensor_desc*>(outputDesc), inputs, outputs, workspace, stream);
    return cerb_trt_corr_enqueue_i64(&f_, reinterpret_cast<const cerb_trt_tensor_desc64*>(inputDesc),
                                     reinterpret_cast<const cerb_trt_tensor_desc64*>(outputDesc), inputs, outputs, workspace, stream);
  }
  size_t getSerializationSize() const noexcept override { return cerb_trt_corr_serialization_size(); }
  void serialize(void* buffer) const noexcept override { cerb_trt_corr_serialize(&f_, buffer); }
  bool supportsFormatCombination(int pos, const nvinfer1::PluginTensorDesc* io, int nbIn, int nbOut) noexcept override {
    if (!kDesc32) return nbIn == 2 && nbOut == 1 && linear_float_or_half(io, pos);
    return cerb_trt_corr_supports_format(pos, reinterpret_cast<const cerb_trt_tensor_desc*>(io), nbIn, nbOut) != 0;
  }
  const char* getPluginType() const noexcept override { return CERB_TRT_CORR_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_CORR_PLUGIN_VERSION; }
  nvinfer1::IPluginV2DynamicExt* clone() const noexcept override { return new CorrelationPlugin(*this); }

 private:
  cerb_trt_corr_fields f_{};
};

// ------------------------------------------------------------------ fused warp + correlation + LeakyReLU ----------
class WarpCorrelationPlugin : public PluginBase {
 public:
  explicit WarpCorrelationPlugin(const nvinfer1::PluginFieldCollection& fc) {
    cerb_trt_warp_corr_default_fields(&f_);
    parse_corr_fields(fc, f_.corr);
    for (int i = 0; i < fc.nbFields; ++i) {
      const char* n = fc.fields[i].name;
      if (!n || !fc.fields[i].data) continue;
      if (!strcmp(n, "warp_mode")) f_.warp_mode = field_int(fc.fields[i]);
      else if (!strcmp(n, "leaky_slope")) f_.leaky_slope = *static_cast<const float*>(fc.fields[i].data);
    }
  }
  WarpCorrelationPlugin(const void* data, size_t length) {
    if (cerb_trt_warp_corr_deserialize(data, length, &f_) != 0) cerb_trt_warp_corr_default_fields(&f_);
  }
  nvinfer1::DimsExprs getOutputDimensions(int, const nvinfer1::DimsExprs* in, int, nvinfer1::IExprBuilder& eb) noexcept override {
    return corr_output_dims(f_.corr, in, eb);
  }
  int enqueue(const nvinfer1::PluginTensorDesc* inputDesc, const nvinfer1::PluginTensorDesc* outputDesc,
              const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept override {
    if (!kDesc32) return CERB_EUNSUPPORTED;   // TensorRT >= 10 descriptors: not wired for the fused node yet
    return cerb_trt_warp_corr_enqueue(&f_.corr, f_.warp_mode, f_.leaky_slope, reinterpret_cast<const cerb_trt_tensor_desc*>(inputDesc),
                                      reinterpret_cast<const cerb_trt_tensor_desc*>(outputDesc), inputs, outputs, workspace, stream);
  }
  size_t getSerializationSize() const noexcept override { return cerb_trt_warp_corr_serialize(&f_, nullptr); }
  void serialize(void* buffer) const noexcept override { cerb_trt_warp_corr_serialize(&f_, buffer); }
  bool supportsFormatCombination(int pos, const nvinfer1::PluginTensorDesc* io, int nbIn, int nbOut) noexcept override {
    return nbIn == 3 && nbOut == 1 && pos >= 0 && pos < 4 && linear_float_or_half(io, pos, 2);
  }
  const char* getPluginType() const noexcept override { return CERB_TRT_WARP_CORR_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_WARP_CORR_PLUGIN_VERSION; }
  nvinfer1::IPluginV2DynamicExt* clone() const noexcept override { return new WarpCorrelationPlugin(*this); }

 private:
  cerb_trt_warp_corr_fields f_{};
};

// ------------------------------------------------------------------ grid sampler ----------
class GridSamplerPlugin : public PluginBase {
 public:
  explicit GridSamplerPlugin(const nvinfer1::PluginFieldCollection& fc) {
    cerb_trt_grid_sampler_default_fields(&f_);   // grid_sampler.cpp:40-42
    for (int i = 0; i < fc.nbFields; ++i) {
      const char* n = fc.fields[i].name;
      if (!n || !fc.fields[i].data) continue;
      if (!strcmp(n, "align_corners")) f_.align_corners = field_int(fc.fields[i]) != 0;
      else if (!strcmp(n, "interpolation_mode")) f_.interpolation_mode = field_int(fc.fields[i]);
      else if (!strcmp(n, "padding_mode")) f_.padding_mode = field_int(fc.fields[i]);
    }
  }
  GridSamplerPlugin(const void* data, size_t length) {
    if (cerb_trt_grid_sampler_deserialize(data, length, &f_) != 0) cerb_trt_grid_sampler_default_fields(&f_);
  }
  // output = (N, C, grid H, grid W)
  nvinfer1::DimsExprs getOutputDimensions(int, const nvinfer1::DimsExprs* in, int, nvinfer1::IExprBuilder&) noexcept override {
    nvinfer1::DimsExprs o;
    o.nbDims = 4;
    o.d[0] = in[0].d[0]; o.d[1] = in[0].d[1]; o.d[2] = in[1].d[1]; o.d[3] = in[1].d[2];
    return o;
  }
  int enqueue(const nvinfer1::PluginTensorDesc* inputDesc, const nvinfer1::PluginTensorDesc* outputDesc,
              const void* const* inputs, void* const* outputs, void* workspace, cudaStream_t stream) noexcept override {
    if (!kDesc32) return CERB_EUNSUPPORTED;
    return cerb_trt_grid_sampler_enqueue(&f_, reinterpret_cast<const cerb_trt_tensor_desc*>(inputDesc),
                                         reinterpret_cast<const cerb_trt_tensor_desc*>(outputDesc), inputs, outputs, workspace, stream);
  }
  size_t getSerializationSize() const noexcept override { return cerb_trt_grid_sampler_serialize(&f_, nullptr); }
  void serialize(void* buffer) const noexcept override { cerb_trt_grid_sampler_serialize(&f_, buffer); }
  bool supportsFormatCombination(int pos, const nvinfer1::PluginTensorDesc* io, int nbIn, int nbOut) noexcept override {
    return nbIn == 2 && nbOut == 1 && pos >= 0 && pos < 3 && linear_float_or_half(io, pos);
  }
  const char* getPluginType() const noexcept override { return CERB_TRT_GRID_SAMPLER_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_GRID_SAMPLER_PLUGIN_VERSION; }
  nvinfer1::IPluginV2DynamicExt* clone() const noexcept override { return new GridSamplerPlugin(*this); }

 private:
  cerb_trt_grid_sampler_fields f_{};
};

// ------------------------------------------------------------------ creators ----------
template <typename Plugin>
class CreatorBase : public nvinfer1::IPluginCreator {
 public:
  const nvinfer1::PluginFieldCollection* getFieldNames() noexcept override { return &fc_; }
  nvinfer1::IPluginV2* createPlugin(const char*, const nvinfer1::PluginFieldCollection* fc) noexcept override {
    static const nvinfer1::PluginFieldCollection none{0, nullptr};
    auto* p = new Plugin(fc ? *fc : none);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  nvinfer1::IPluginV2* deserializePlugin(const char*, const void* data, size_t len) noexcept override {
    auto* p = new Plugin(data, len);
    p->setPluginNamespace(ns_.c_str());
    return p;
  }
  void setPluginNamespace(const char* ns) noexcept override { ns_ = ns ? ns : ""; }
  const char* getPluginNamespace() const noexcept override { return ns_.c_str(); }

 protected:
  void add(const char* name, nvinfer1::PluginFieldType t) {
    attrs_.emplace_back(nvinfer1::PluginField(name, nullptr, t, 1));
    fc_.nbFields = (int)attrs_.size();
    fc_.fields = attrs_.data();
  }
  nvinfer1::PluginFieldCollection fc_{};
  std::vector<nvinfer1::PluginField> attrs_;
  std::string ns_;
};

static const char* const kCorrFieldNames[6] = {"pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply"};

class CorrelationPluginCreator : public CreatorBase<CorrelationPlugin> {   // correlation.cpp:268-273
 public:
  CorrelationPluginCreator() { attrs_.reserve(8); for (const char* n : kCorrFieldNames) add(n, nvinfer1::PluginFieldType::kINT32); }
  const char* getPluginName() const noexcept override { return CERB_TRT_CORR_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_CORR_PLUGIN_VERSION; }
};
class WarpCorrelationPluginCreator : public CreatorBase<WarpCorrelationPlugin> {
 public:
  WarpCorrelationPluginCreator() {
    attrs_.reserve(8);
    for (const char* n : kCorrFieldNames) add(n, nvinfer1::PluginFieldType::kINT32);
    add("warp_mode", nvinfer1::PluginFieldType::kINT32);
    add("leaky_slope", nvinfer1::PluginFieldType::kFLOAT32);
  }
  const char* getPluginName() const noexcept override { return CERB_TRT_WARP_CORR_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_WARP_CORR_PLUGIN_VERSION; }
};
class GridSamplerPluginCreator : public CreatorBase<GridSamplerPlugin> {   // grid_sampler.cpp:196-198
 public:
  GridSamplerPluginCreator() {
    attrs_.reserve(4);
    add("align_corners", nvinfer1::PluginFieldType::kINT32);
    add("interpolation_mode", nvinfer1::PluginFieldType::kINT32);
    add("padding_mode", nvinfer1::PluginFieldType::kINT32);
  }
  const char* getPluginName() const noexcept override { return CERB_TRT_GRID_SAMPLER_PLUGIN_TYPE; }
  const char* getPluginVersion() const noexcept override { return CERB_TRT_GRID_SAMPLER_PLUGIN_VERSION; }
};

REGISTER_TENSORRT_PLUGIN(CorrelationPluginCreator);        // correlation.hpp:108
REGISTER_TENSORRT_PLUGIN(WarpCorrelationPluginCreator);
REGISTER_TENSORRT_PLUGIN(GridSamplerPluginCreator);        // grid_sampler.hpp:112

}  // namespace cerb_trt
#endif  // CERB_HAVE_TENSORRT
