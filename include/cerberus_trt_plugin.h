/*
 * cerberus_trt_plugin.h -- C ABI shaped like the reference's TensorRT correlation plugin.
 *
 * TensorRT is not part of this build; these entry points are what a CorrelationPlugin
 * (runtime/cerberus_net/trt_plugins/correlation.hpp:10-72) forwards to.  The descriptor
 * structs are layout-compatible with TensorRT 7/8's nvinfer1::Dims / PluginTensorDesc
 * (int32 nbDims + int32 d[8]; DataType; TensorFormat; float scale), so the plugin shim in
 * cerberusnet_b200/csrc/trt_plugin_shim.cpp can reinterpret_cast its arguments.
 * (TensorRT >= 10 widens d[] to int64: use the *_i64 variant.)
 */
#ifndef CERBERUS_TRT_PLUGIN_H_
#define CERBERUS_TRT_PLUGIN_H_

#include "cerberus_costvolume.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CERB_TRT_MAX_DIMS 8
enum { CERB_TRT_FLOAT = 0, CERB_TRT_HALF = 1 }; /* nvinfer1::DataType::kFLOAT / kHALF */
enum { CERB_TRT_LINEAR = 0 };                   /* nvinfer1::TensorFormat::kLINEAR */

typedef struct cerb_trt_dims { int32_t nbDims; int32_t d[CERB_TRT_MAX_DIMS]; } cerb_trt_dims;
typedef struct cerb_trt_tensor_desc {
  cerb_trt_dims dims;
  int32_t type;   /* CERB_TRT_FLOAT / CERB_TRT_HALF */
  int32_t format; /* CERB_TRT_LINEAR */
  float scale;
} cerb_trt_tensor_desc;

typedef struct cerb_trt_dims64 { int32_t nbDims; int64_t d[CERB_TRT_MAX_DIMS]; } cerb_trt_dims64;
typedef struct cerb_trt_tensor_desc64 {
  cerb_trt_dims64 dims;
  int32_t type;
  int32_t format;
  float scale;
} cerb_trt_tensor_desc64;

/* The plugin's six kINT32 fields, in serialisation order: exactly the 24 bytes
 * CorrelationPlugin::serialize writes (trt_plugins/correlation.cpp:78-105).
 * Defaults when no fields are given: 4,1,4,1,1,1 (correlation.cpp:54-62). */
typedef struct cerb_trt_corr_fields {
  int32_t pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply;
} cerb_trt_corr_fields;

#define CERB_TRT_CORR_PLUGIN_TYPE "correlation" /* correlation.cpp:11 */
#define CERB_TRT_CORR_PLUGIN_VERSION "1"        /* correlation.cpp:10 */

CERB_API void cerb_trt_corr_default_fields(cerb_trt_corr_fields* f);

/* (de)serialisation: replaces CorrelationPlugin(const void*, size_t) / serialize /
 * getSerializationSize (correlation.cpp:65-105).  Returns bytes written / consumed (24). */
CERB_API size_t cerb_trt_corr_serialization_size(void);
CERB_API size_t cerb_trt_corr_serialize(const cerb_trt_corr_fields* f, void* buffer);
CERB_API int cerb_trt_corr_deserialize(const void* data, size_t length, cerb_trt_corr_fields* f);

/* replaces CorrelationPlugin::getOutputDimensions (correlation.cpp:178-205) */
CERB_API int cerb_trt_corr_output_dims(const cerb_trt_corr_fields* f, const cerb_trt_dims* in0,
                                       cerb_trt_dims* out);

/* replaces CorrelationPlugin::supportsFormatCombination (correlation.cpp:144-176);
 * returns 1 / 0 */
CERB_API int cerb_trt_corr_supports_format(int pos, const cerb_trt_tensor_desc* in_out, int nb_inputs,
                                           int nb_outputs);

/* replaces CorrelationPlugin::getWorkspaceSize (correlation.cpp:114-136): always 0 here
 * (the reference asks for 2*N*C*(H+2p)*(W+2p)*sizeof(T) and then over-runs it,
 * trt_plugins/correlation.cu:112,137). */
CERB_API size_t cerb_trt_corr_workspace_size(const cerb_trt_corr_fields* f, const cerb_trt_tensor_desc* inputs,
                                             int nb_inputs, const cerb_trt_tensor_desc* outputs, int nb_outputs);

/* replaces CorrelationPlugin::enqueue (correlation.hpp:31-32, correlation.cu:94-166).
 * Same argument order and meaning; returns int(cudaError_t) like `return cudaGetLastError()`
 * (0 = ok), or a negative CERB_E* for descriptor errors instead of throwing
 * (correlation.cu:159-162).  inputs: {x1, x2}; outputs: {cost volume}; workspace unused. */
CERB_API int cerb_trt_corr_enqueue(const cerb_trt_corr_fields* f, const cerb_trt_tensor_desc* input_desc,
                                   const cerb_trt_tensor_desc* output_desc, const void* const* inputs,
                                   void* const* outputs, void* workspace, cerb_stream_t stream);
CERB_API int cerb_trt_corr_enqueue_i64(const cerb_trt_corr_fields* f, const cerb_trt_tensor_desc64* input_desc,
                                       const cerb_trt_tensor_desc64* output_desc, const void* const* inputs,
                                       void* const* outputs, void* workspace, cerb_stream_t stream);

/* Fused node that replaces, per pyramid level of the exported graph, 2x ScatterND + Transpose
 * + grid_sampler + correlation + LeakyRelu: inputs {im1, im2, flow(fp32)},
 * output {activated cost volume}.  warp_mode defaults to CERB_WARP_TRT in the shim. */
CERB_API int cerb_trt_warp_corr_enqueue(const cerb_trt_corr_fields* f, int32_t warp_mode, float leaky_slope,
                                        const cerb_trt_tensor_desc* input_desc,
                                        const cerb_trt_tensor_desc* output_desc, const void* const* inputs,
                                        void* const* outputs, void* workspace, cerb_stream_t stream);

/* ---- fused warp_correlation node: plugin type "warp_correlation", version "1" (new; replaces the five nodes above).
 * Fields: the six correlation ints, then warp_mode (kINT32, default CERB_WARP_TRT) and leaky_slope (kFLOAT32,
 * default 0.1); serialised in that order, 32 bytes. */
typedef struct cerb_trt_warp_corr_fields {
  cerb_trt_corr_fields corr;
  int32_t warp_mode;
  float leaky_slope;
} cerb_trt_warp_corr_fields;
#define CERB_TRT_WARP_CORR_PLUGIN_TYPE "warp_correlation"
#define CERB_TRT_WARP_CORR_PLUGIN_VERSION "1"
CERB_API void cerb_trt_warp_corr_default_fields(cerb_trt_warp_corr_fields* f);
CERB_API size_t cerb_trt_warp_corr_serialize(const cerb_trt_warp_corr_fields* f, void* buffer);   /* 32 bytes, NULL buffer = size query */
CERB_API int cerb_trt_warp_corr_deserialize(const void* data, size_t length, cerb_trt_warp_corr_fields* f);

/* ---- grid_sampler plugin: type "grid_sampler", version "1" (trt_plugins/grid_sampler.cpp:9-10).
 * Fields align_corners, interpolation_mode, padding_mode, all kINT32 (grid_sampler.cpp:196-198); defaults false /
 * Bilinear / Border (:40-42); serialisation = 1-byte bool + two ints = 9 bytes (:57-74).  The enumerations are
 * CERB_GRID_* (grid_sampler.hpp:14-15). */
typedef struct cerb_trt_grid_sampler_fields {
  int32_t align_corners, interpolation_mode, padding_mode;
} cerb_trt_grid_sampler_fields;
#define CERB_TRT_GRID_SAMPLER_PLUGIN_TYPE "grid_sampler"
#define CERB_TRT_GRID_SAMPLER_PLUGIN_VERSION "1"
CERB_API void cerb_trt_grid_sampler_default_fields(cerb_trt_grid_sampler_fields* f);
CERB_API size_t cerb_trt_grid_sampler_serialize(const cerb_trt_grid_sampler_fields* f, void* buffer);   /* 9 bytes, NULL buffer = size query */
CERB_API int cerb_trt_grid_sampler_deserialize(const void* data, size_t length, cerb_trt_grid_sampler_fields* f);
/* replaces GridSamplerPlugin::enqueue (grid_sampler.cu:238-271): inputs {input (N,C,H,W), grid (N,H,W,2)} of one
 * type (kFLOAT or kHALF), output (N,C,H,W); the plugin's own un-normalise / rounding (CERB_GRID_CONV_TRT). */
CERB_API int cerb_trt_grid_sampler_enqueue(const cerb_trt_grid_sampler_fields* f, const cerb_trt_tensor_desc* input_desc,
                                           const cerb_trt_tensor_desc* output_desc, const void* const* inputs,
                                           void* const* outputs, void* workspace, cerb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CERBERUS_TRT_PLUGIN_H_ */
