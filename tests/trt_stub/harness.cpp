// harness.cpp -- drives cerberusnet_b200/csrc/trt_plugin_shim.cpp through tests/trt_stub/NvInfer.h the way TensorRT's
// builder / runtime would: creator lookup by (type, version), createPlugin from a PluginFieldCollection,
// getOutputDimensions through an IExprBuilder, serialize -> deserializePlugin -> serialize, clone, format support,
// and (with --gpu) enqueue on device buffers compared bit for bit with the direct C-ABI call.
// Test infrastructure; built and run by tests/test_trt_shim.py.
#include <NvInfer.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../include/cerberus_trt_plugin.h"

using namespace nvinfer1;

static int g_fail = 0;
#define CHECK(cond) do { if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_fail; } } while (0)

// ---- a concrete expression builder (constants only: shapes are static here)
struct Expr : IDimensionExpr {
  int32_t v;
  explicit Expr(int32_t x) : v(x) {}
  bool isConstant() const noexcept override { return true; }
  int32_t getConstantValue() const noexcept override { return v; }
};
struct Builder : IExprBuilder {
  std::vector<std::unique_ptr<Expr>> pool;
  const IDimensionExpr* constant(int32_t v) noexcept override { pool.emplace_back(new Expr(v)); return pool.back().get(); }
  const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& a, const IDimensionExpr& b) noexcept override {
    const int32_t x = a.getConstantValue(), y = b.getConstantValue();
    int32_t r = 0;
    switch (op) {
      case DimensionOperation::kSUM: r = x + y; break;
      case DimensionOperation::kSUB: r = x - y; break;
      case DimensionOperation::kPROD: r = x * y; break;
      case DimensionOperation::kCEIL_DIV: r = (x + y - 1) / y; break;
      case DimensionOperation::kFLOOR_DIV: r = x / y; break;
      case DimensionOperation::kMAX: r = x > y ? x : y; break;
      case DimensionOperation::kMIN: r = x < y ? x : y; break;
      default: r = 0;
    }
    return constant(r);
  }
};

static DimsExprs dims_of(Builder& b, std::initializer_list<int> d) {
  DimsExprs e;
  e.nbDims = (int)d.size();
  int i = 0;
  for (int v : d) e.d[i++] = b.constant(v);
  return e;
}
static PluginTensorDesc desc(std::initializer_list<int> d, DataType t) {
  PluginTensorDesc p{};
  p.dims.nbDims = (int)d.size();
  int i = 0;
  for (int v : d) p.dims.d[i++] = v;
  p.type = t;
  p.format = TensorFormat::kLINEAR;
  p.scale = 1.f;
  return p;
}

static IPluginV2DynamicExt* roundtrip(IPluginCreator* c, IPluginV2DynamicExt* p, size_t expect_bytes) {
  CHECK(p->getSerializationSize() == expect_bytes);
  std::vector<unsigned char> buf(p->getSerializationSize()), buf2(p->getSerializationSize());
  p->serialize(buf.data());
  auto* q = static_cast<IPluginV2DynamicExt*>(c->deserializePlugin("x", buf.data(), buf.size()));
  CHECK(q != nullptr);
  CHECK(q->getSerializationSize() == expect_bytes);
  q->serialize(buf2.data());
  CHECK(buf == buf2);
  auto* r = q->clone();
  std::vector<unsigned char> buf3(r->getSerializationSize());
  r->serialize(buf3.data());
  CHECK(buf == buf3);
  r->destroy();
  return q;
}

static void cpu_checks() {
  auto* reg = getPluginRegistry();
  IPluginCreator* cc = reg->getPluginCreator("correlation", "1");
  IPluginCreator* wc = reg->getPluginCreator("warp_correlation", "1");
  IPluginCreator* gc = reg->getPluginCreator("grid_sampler", "1");
  CHECK(cc && wc && gc);
  if (!cc || !wc || !gc) return;
  CHECK(reg->getPluginCreator("correlation", "2") == nullptr);

  // field names the exporter relies on (onnx_export.py:21-23,27-28; correlation.cpp:268-273; grid_sampler.cpp:196-198)
  const PluginFieldCollection* fn = cc->getFieldNames();
  CHECK(fn->nbFields == 6);
  const char* names[6] = {"pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply"};
  for (int i = 0; i < 6 && i < fn->nbFields; ++i) { CHECK(!strcmp(fn->fields[i].name, names[i])); CHECK(fn->fields[i].type == PluginFieldType::kINT32); }
  CHECK(wc->getFieldNames()->nbFields == 8);
  CHECK(gc->getFieldNames()->nbFields == 3);

  Builder b;
  {   // ---- correlation: defaults when no fields are given are 4,1,4,1,1,1 (correlation.cpp:54-62)
    PluginFieldCollection none{0, nullptr};
    auto* p = static_cast<IPluginV2DynamicExt*>(cc->createPlugin("corr", &none));
    CHECK(!strcmp(p->getPluginType(), "correlation") && !strcmp(p->getPluginVersion(), "1"));
    CHECK(p->getNbOutputs() == 1);
    DimsExprs in[2] = {dims_of(b, {2, 64, 64, 128}), dims_of(b, {2, 64, 64, 128})};
    DimsExprs o = p->getOutputDimensions(0, in, 2, b);
    CHECK(o.nbDims == 4 && o.d[0]->getConstantValue() == 2 && o.d[1]->getConstantValue() == 81 &&
          o.d[2]->getConstantValue() == 64 && o.d[3]->getConstantValue() == 128);
    int ser[6];
    CHECK(p->getSerializationSize() == 24);
    p->serialize(ser);
    const int expect[6] = {4, 1, 4, 1, 1, 1};
    for (int i = 0; i < 6; ++i) CHECK(ser[i] == expect[i]);
    PluginTensorDesc io[3] = {desc({2, 64, 64, 128}, DataType::kFLOAT), desc({2, 64, 64, 128}, DataType::kFLOAT), desc({2, 81, 64, 128}, DataType::kFLOAT)};
    for (int pos = 0; pos < 3; ++pos) CHECK(p->supportsFormatCombination(pos, io, 2, 1));
    io[1].type = DataType::kHALF;
    CHECK(!p->supportsFormatCombination(1, io, 2, 1));
    io[1].type = DataType::kINT8; io[0].type = DataType::kINT8;
    CHECK(!p->supportsFormatCombination(0, io, 2, 1));
    CHECK(p->getWorkspaceSize(io, 2, io + 2, 1) == 0);
    CHECK(p->initialize() == 0);
    DataType t = DataType::kHALF;
    CHECK(p->getOutputDataType(0, &t, 2) == DataType::kHALF);
    p->setPluginNamespace("ns");
    CHECK(!strcmp(p->getPluginNamespace(), "ns"));
    p->destroy();
  }
  {   // ---- correlation with explicit fields: the reference's __main__ case pad 4, md 10 (correlation.py:87) -> (H-12) x (W-12), 441 planes
    int v[6] = {4, 1, 10, 1, 1, 1};
    PluginField f[6];
    for (int i = 0; i < 6; ++i) f[i] = PluginField(names[i], &v[i], PluginFieldType::kINT32, 1);
    PluginFieldCollection fc{6, f};
    auto* p = static_cast<IPluginV2DynamicExt*>(cc->createPlugin("corr", &fc));
    DimsExprs in[2] = {dims_of(b, {1, 16, 40, 52}), dims_of(b, {1, 16, 40, 52})};
    DimsExprs o = p->getOutputDimensions(0, in, 2, b);
    CHECK(o.d[1]->getConstantValue() == 441 && o.d[2]->getConstantValue() == 28 && o.d[3]->getConstantValue() == 40);
    auto* q = roundtrip(cc, p, 24);
    int ser[6];
    q->serialize(ser);
    for (int i = 0; i < 6; ++i) CHECK(ser[i] == v[i]);
    q->destroy();
    p->destroy();
  }
  {   // ---- fused node
    int md = 4, mode = CERB_WARP_TORCH;
    float slope = 0.2f;
    PluginField f[3] = {PluginField("max_displacement", &md, PluginFieldType::kINT32, 1), PluginField("warp_mode", &mode, PluginFieldType::kINT32, 1),
                        PluginField("leaky_slope", &slope, PluginFieldType::kFLOAT32, 1)};
    PluginFieldCollection fc{3, f};
    auto* p = static_cast<IPluginV2DynamicExt*>(wc->createPlugin("wc", &fc));
    CHECK(!strcmp(p->getPluginType(), "warp_correlation"));
    DimsExprs in[3] = {dims_of(b, {1, 32, 128, 256}), dims_of(b, {1, 32, 128, 256}), dims_of(b, {1, 2, 128, 256})};
    DimsExprs o = p->getOutputDimensions(0, in, 3, b);
    CHECK(o.d[1]->getConstantValue() == 81 && o.d[2]->getConstantValue() == 128 && o.d[3]->getConstantValue() == 256);
    auto* q = roundtrip(wc, p, 32);
    cerb_trt_warp_corr_fields got;
    q->serialize(&got);
    CHECK(got.warp_mode == CERB_WARP_TORCH && got.leaky_slope == 0.2f && got.corr.pad_size == 4 && got.corr.stride2 == 1);
    PluginTensorDesc io[4] = {desc({1, 32, 128, 256}, DataType::kHALF), desc({1, 32, 128, 256}, DataType::kHALF),
                              desc({1, 2, 128, 256}, DataType::kFLOAT), desc({1, 81, 128, 256}, DataType::kHALF)};
    for (int pos = 0; pos < 4; ++pos) CHECK(p->supportsFormatCombination(pos, io, 3, 1));
    io[2].type = DataType::kHALF;
    CHECK(!p->supportsFormatCombination(2, io, 3, 1));   // the flow stays fp32
    q->destroy();
    p->destroy();
  }
  {   // ---- grid sampler: defaults false / Bilinear / Border (grid_sampler.cpp:40-42), 9-byte serialisation (:57-74)
    PluginFieldCollection none{0, nullptr};
    auto* p = static_cast<IPluginV2DynamicExt*>(gc->createPlugin("gs", &none));
    unsigned char ser[9];
    CHECK(p->getSerializationSize() == 9);
    p->serialize(ser);
    int im, pm;
    memcpy(&im, ser + 1, 4); memcpy(&pm, ser + 5, 4);
    CHECK(ser[0] == 0 && im == CERB_GRID_BILINEAR && pm == CERB_GRID_PAD_BORDER);
    DimsExprs in[2] = {dims_of(b, {2, 48, 32, 64}), dims_of(b, {2, 32, 64, 2})};
    DimsExprs o = p->getOutputDimensions(0, in, 2, b);
    CHECK(o.d[0]->getConstantValue() == 2 && o.d[1]->getConstantValue() == 48 && o.d[2]->getConstantValue() == 32 && o.d[3]->getConstantValue() == 64);
    p->destroy();
    int ac = 1, imode = CERB_GRID_NEAREST, pmode = CERB_GRID_PAD_REFLECTION;
    PluginField f[3] = {PluginField("align_corners", &ac, PluginFieldType::kINT32, 1), PluginField("interpolation_mode", &imode, PluginFieldType::kINT32, 1),
                        PluginField("padding_mode", &pmode, PluginFieldType::kINT32, 1)};
    PluginFieldCollection fc{3, f};
    auto* p2 = static_cast<IPluginV2DynamicExt*>(gc->createPlugin("gs", &fc));
    auto* q = roundtrip(gc, p2, 9);
    q->serialize(ser);
    memcpy(&im, ser + 1, 4); memcpy(&pm, ser + 5, 4);
    CHECK(ser[0] == 1 && im == CERB_GRID_NEAREST && pm == CERB_GRID_PAD_REFLECTION);
    q->destroy();
    p2->destroy();
  }
}

// ------------------------------------------------------------------ GPU: enqueue == the C ABI call, bit for bit
#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL cuda %s at %d\n", cudaGetErrorString(e_), __LINE__); ++g_fail; return; } } while (0)

static void fill(std::vector<float>& v, unsigned seed, float scale) {
  for (auto& x : v) { seed = seed * 1664525u + 1013904223u; x = scale * ((float)((seed >> 8) & 0xffff) / 32768.f - 1.f); }
}

static void gpu_checks() {
  auto* reg = getPluginRegistry();
  const int N = 2, C = 24, H = 40, W = 64;
  const size_t ne = (size_t)N * C * H * W, nf = (size_t)N * 2 * H * W, no = (size_t)N * 81 * H * W;
  std::vector<float> h1(ne), h2(ne), hf(nf), hg(nf);
  fill(h1, 1, 1.f); fill(h2, 2, 1.f); fill(hf, 3, 5.f); fill(hg, 4, 1.2f);
  float *d1, *d2, *df, *dg, *oa, *ob;
  CUDA_OK(cudaMalloc(&d1, ne * 4)); CUDA_OK(cudaMalloc(&d2, ne * 4)); CUDA_OK(cudaMalloc(&df, nf * 4)); CUDA_OK(cudaMalloc(&dg, nf * 4));
  CUDA_OK(cudaMalloc(&oa, no * 4)); CUDA_OK(cudaMalloc(&ob, no * 4));
  CUDA_OK(cudaMemcpy(d1, h1.data(), ne * 4, cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(d2, h2.data(), ne * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(df, hf.data(), nf * 4, cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(dg, hg.data(), nf * 4, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CUDA_OK(cudaStreamCreate(&st));
  std::vector<float> ra(no), rb(no);
  PluginFieldCollection none{0, nullptr};
  cerb_corr_params p{};
  p.batch = N; p.channels = C; p.height = H; p.width = W; p.pad_size = 4; p.kernel_size = 1; p.max_displacement = 4;
  p.stride1 = p.stride2 = 1; p.corr_multiply = 1; p.dtype = CERB_F32; p.warp_mode = CERB_WARP_TRT; p.leaky_slope = NAN;

  {   // correlation plugin == cerb_warp_corr_forward(no flow, no activation)
    auto* pl = static_cast<IPluginV2DynamicExt*>(reg->getPluginCreator("correlation", "1")->createPlugin("c", &none));
    PluginTensorDesc in[2] = {desc({N, C, H, W}, DataType::kFLOAT), desc({N, C, H, W}, DataType::kFLOAT)};
    PluginTensorDesc out = desc({N, 81, H, W}, DataType::kFLOAT);
    const void* ins[2] = {d1, d2};
    void* outs[1] = {oa};
    CHECK(pl->enqueue(in, &out, ins, outs, nullptr, st) == 0);
    CHECK(cerb_warp_corr_forward(&p, d1, d2, nullptr, ob, st) == 0);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaMemcpy(ra.data(), oa, no * 4, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(rb.data(), ob, no * 4, cudaMemcpyDeviceToHost));
    CHECK(memcmp(ra.data(), rb.data(), no * 4) == 0);
    out.dims.d[2] = H - 1;   // wrong output extent: an error code, not an abort
    CHECK(pl->enqueue(in, &out, ins, outs, nullptr, st) == CERB_ESHAPE);
    pl->destroy();
  }
  {   // fused node == cerb_warp_corr_forward(flow, TRT warp, slope 0.1)
    auto* pl = static_cast<IPluginV2DynamicExt*>(reg->getPluginCreator("warp_correlation", "1")->createPlugin("w", &none));
    PluginTensorDesc in[3] = {desc({N, C, H, W}, DataType::kFLOAT), desc({N, C, H, W}, DataType::kFLOAT), desc({N, 2, H, W}, DataType::kFLOAT)};
    PluginTensorDesc out = desc({N, 81, H, W}, DataType::kFLOAT);
    const void* ins[3] = {d1, d2, df};
    void* outs[1] = {oa};
    CHECK(pl->enqueue(in, &out, ins, outs, nullptr, st) == 0);
    p.leaky_slope = 0.1f;
    CHECK(cerb_warp_corr_forward(&p, d1, d2, df, ob, st) == 0);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaMemcpy(ra.data(), oa, no * 4, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(rb.data(), ob, no * 4, cudaMemcpyDeviceToHost));
    CHECK(memcmp(ra.data(), rb.data(), no * 4) == 0);
    pl->destroy();
  }
  {   // grid sampler plugin == cerb_grid_sample_forward(TRT convention)
    auto* pl = static_cast<IPluginV2DynamicExt*>(reg->getPluginCreator("grid_sampler", "1")->createPlugin("g", &none));
    PluginTensorDesc in[2] = {desc({N, C, H, W}, DataType::kFLOAT), desc({N, H, W, 2}, DataType::kFLOAT)};
    PluginTensorDesc out = desc({N, C, H, W}, DataType::kFLOAT);
    const void* ins[2] = {d1, dg};
    void* outs[1] = {oa};
    CHECK(pl->enqueue(in, &out, ins, outs, nullptr, st) == 0);
    CHECK(cerb_grid_sample_forward(d1, dg, ob, N, C, H, W, H, W, CERB_F32, CERB_GRID_BILINEAR, CERB_GRID_PAD_BORDER, 0, CERB_GRID_CONV_TRT, st) == 0);
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaMemcpy(ra.data(), oa, ne * 4, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(rb.data(), ob, ne * 4, cudaMemcpyDeviceToHost));
    CHECK(memcmp(ra.data(), rb.data(), ne * 4) == 0);
    pl->destroy();
  }
  cudaFree(d1); cudaFree(d2); cudaFree(df); cudaFree(dg); cudaFree(oa); cudaFree(ob);
  cudaStreamDestroy(st);
}

int main(int argc, char** argv) {
  cpu_checks();
  if (argc > 1 && !strcmp(argv[1], "--gpu")) gpu_checks();
  if (g_fail == 0) printf("OK\n");
  return g_fail == 0 ? 0 : 1;
}
