/*
 * idn_host.h -- flat C view of the host-side mirror of idencomp's API (libidn_host.so).
 *
 * The host mirror itself is C++ (idencomp_b200/csrc/host/, the .hpp files: Model, ModelProvider, IdnCompressor,
 * IdnDecompressor -- same names, arguments and error behaviour as the reference's Rust types) and calls
 * only the C-ABI of include/idn_gpu.h.  This header exists so that the Python test/bench harness (ctypes)
 * and other FFI users can drive those classes; every function cites the reference item it stands for.
 * Status codes are the IDN_* codes of idn_gpu.h.  idn_host_last_error() is per thread.
 */
#ifndef IDN_HOST_H
#define IDN_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "idn_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct idn_host_model idn_host_model;

const char *idn_host_last_error(void);

/* ---- Model (model.rs:176-259, model_serializer.rs:90-114) ---------------------------------------- */
/* SerializableModel::read_model: parses the msgpack file, recomputes the SHA3-256 identifier and fails
 * with IDN_E_SERIALIZE if it differs from the stored one (model_serializer.rs:111-114). */
int32_t idn_host_model_load(const char *path, idn_host_model **out);
int32_t idn_host_model_from_bytes(const uint8_t *data, size_t n, idn_host_model **out);
/* Model::with_model_and_spec_type from raw contexts: probs[n_ctx][nsym] f32, spec lists given as
 * (spec_keys[i] -> context spec_ctx[i]) pairs (used by tests to build the reference's toy models). */
int32_t idn_host_model_new(int32_t model_type, const char *spec_name, uint32_t n_ctx, const float *probs,
                           const uint32_t *spec_keys, const uint32_t *spec_ctx, uint64_t n_specs,
                           idn_host_model **out);
int32_t idn_host_model_empty(int32_t model_type, idn_host_model **out); /* Model::empty, model.rs:251-259 */
void idn_host_model_free(idn_host_model *m);
int32_t idn_host_model_type(const idn_host_model *m);          /* IDN_MODEL_ACID / IDN_MODEL_QSCORE */
uint32_t idn_host_model_len(const idn_host_model *m);          /* number of contexts */
const char *idn_host_model_spec_name(const idn_host_model *m); /* e.g. "light_ao8_qo0_pb0_qm1" */
void idn_host_model_identifier(const idn_host_model *m, uint8_t out[32]);
/* integer tables: cum [(len+1)][nsym+1] u16 (row 0 = dummy context); returns the number of u16 written
 * (call with out == NULL to size) */
uint64_t idn_host_model_cum_table(const idn_host_model *m, uint16_t *out, uint64_t cap);
/* RansEncModel/RansDecModel::from_model (sequence_compressor.rs:21-48,175-201) on the device:
 * quantises every context (context.rs:346-394) and uploads the tables through idn_gpu_model_upload */
int32_t idn_host_model_upload(idn_gpu_ctx *ctx, const idn_host_model *m, idn_model_t *handle);

/* Context::as_integer_cum_freqs (context.rs:346-371) -- exposed for the known-answer tests */
int32_t idn_host_quantise(const float *probs, uint32_t nsym, uint32_t scale_bits, uint32_t *cum_out);

#ifdef __cplusplus
}
#endif
#endif /* IDN_HOST_H */
