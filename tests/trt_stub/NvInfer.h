// NvInfer.h -- MINIMAL STAND-IN for TensorRT's public header, test infrastructure only.
//
// TensorRT is not in this image.  This stub declares, from the public TensorRT 8 plugin API, exactly the types and
// virtual methods that the reference's plugins implement (runtime/cerberus_net/trt_plugins/correlation.hpp:10-108,
// grid_sampler.hpp:21-112) and that cerberusnet_b200/csrc/trt_plugin_shim.cpp overrides, so the shim is compiled and
// its plugin classes are exercised (tests/test_trt_shim.py) instead of being dead code.  It is NOT TensorRT: there is no
// builder, engine or runtime here, and the registry is a 16-entry table.
#pragma once
#include <cuda_runtime_api.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace nvinfer1 {

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
enum class TensorFormat : int32_t { kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3 };

struct Dims {
  static constexpr int32_t MAX_DIMS = 8;
  int32_t nbDims;
  int32_t d[MAX_DIMS];
};
struct PluginTensorDesc {
  Dims dims;
  DataType type;
  TensorFormat format;
  float scale;
};
struct DynamicPluginTensorDesc {
  PluginTensorDesc desc;
  Dims min;
  Dims max;
};

enum class DimensionOperation : int32_t { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8 };
class IDimensionExpr {
 public:
  virtual bool isConstant() const noexcept = 0;
  virtual int32_t getConstantValue() const noexcept = 0;
 protected:
  virtual ~IDimensionExpr() = default;
};
struct DimsExprs {
  int32_t nbDims;
  const IDimensionExpr* d[Dims::MAX_DIMS];
};
class IExprBuilder {
 public:
  virtual const IDimensionExpr* constant(int32_t value) noexcept = 0;
  virtual const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& first, const IDimensionExpr& second) noexcept = 0;
 protected:
  virtual ~IExprBuilder() = default;
};

enum class PluginFieldType : int32_t { kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8 };
struct PluginField {
  const char* name;
  const void* data;
  PluginFieldType type;
  int32_t length;
  PluginField(const char* n = nullptr, const void* d = nullptr, PluginFieldType t = PluginFieldType::kUNKNOWN, int32_t l = 0)
      : name(n), data(d), type(t), length(l) {}
};
struct PluginFieldCollection {
  int32_t nbFields;
  const PluginField* fields;
};

class IPluginV2 {
 public:
  virtual int32_t getNbOutputs() const noexcept = 0;
  virtual int32_t initialize() noexcept = 0;
  virtual void terminate() noexcept = 0;
  virtual size_t getSerializationSize() const noexcept = 0;
  virtual void serialize(void* buffer) const noexcept = 0;
  virtual const char* getPluginType() const noexcept = 0;
  virtual const char* getPluginVersion() const noexcept = 0;
  virtual void destroy() noexcept = 0;
  virtual void setPluginNamespace(const char* ns) noexcept = 0;
  virtual const char* getPluginNamespace() const noexcept = 0;
  virtual ~IPluginV2() = default;
};
class IPluginV2Ext : public IPluginV2 {
 public:
  virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t nbInputs) const noexcept = 0;
};
class IPluginV2DynamicExt : public IPluginV2Ext {
 public:
  virtual IPluginV2DynamicExt* clone() const noexcept = 0;
  virtual DimsExprs getOutputDimensions(int32_t outputIndex, const DimsExprs* inputs, int32_t nbInputs, IExprBuilder& exprBuilder) noexcept = 0;
  virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs, int32_t nbOutputs) noexcept = 0;
  virtual void configurePlugin(const DynamicPluginTensorDesc* in, int32_t nbInputs, const DynamicPluginTensorDesc* out, int32_t nbOutputs) noexcept = 0;
  virtual size_t getWorkspaceSize(const PluginTensorDesc* inputs, int32_t nbInputs, const PluginTensorDesc* outputs, int32_t nbOutputs) const noexcept = 0;
  virtual int32_t enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc, const void* const* inputs,
                          void* const* outputs, void* workspace, cudaStream_t stream) noexcept = 0;
};

class IPluginCreator {
 public:
  virtual const char* getPluginName() const noexcept = 0;
  virtual const char* getPluginVersion() const noexcept = 0;
  virtual const PluginFieldCollection* getFieldNames() noexcept = 0;
  virtual IPluginV2* createPlugin(const char* name, const PluginFieldCollection* fc) noexcept = 0;
  virtual IPluginV2* deserializePlugin(const char* name, const void* serialData, size_t serialLength) noexcept = 0;
  virtual void setPluginNamespace(const char* ns) noexcept = 0;
  virtual const char* getPluginNamespace() const noexcept = 0;
  virtual ~IPluginCreator() = default;
};

// ---- a 16-entry creator table standing in for getPluginRegistry()
struct StubRegistry {
  IPluginCreator* creators[16];
  int n;
  bool registerCreator(IPluginCreator& c, const char*) noexcept {
    if (n >= 16) return false;
    creators[n++] = &c;
    return true;
  }
  IPluginCreator* getPluginCreator(const char* type, const char* version, const char* = "") noexcept {
    for (int i = 0; i < n; ++i)
      if (!strcmp(creators[i]->getPluginName(), type) && !strcmp(creators[i]->getPluginVersion(), version)) return creators[i];
    return nullptr;
  }
};
inline StubRegistry* getPluginRegistry() {
  static StubRegistry r{};
  return &r;
}
template <typename T>
class PluginRegistrar {
 public:
  PluginRegistrar() { getPluginRegistry()->registerCreator(instance, ""); }
 private:
  T instance{};
};

}  // namespace nvinfer1

#define REGISTER_TENSORRT_PLUGIN(name) static nvinfer1::PluginRegistrar<name> pluginRegistrar##name {}
