/*
 * idn_oracle.h -- CPU restatement of idencomp's rANS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and there
 * only as the checker / the timed CPU baseline.  The product (idencomp_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against every golden
 * vector the reference's tests hold for the path: samples/1M.idn <-> samples/1M.fastq
 * (decode AND bit-exact re-encode), the quantiser / context-spec / rANS known-answer constants, the SHA3
 * model identifiers.  Unpinned corners (no reference golden exists): Brotli name slices, clustering RNG.
 *
 * The Rust reference cannot be compiled in this image (no rustc/cargo, crates.io dependencies), so there
 * is no oracle/_ref build; see DESIGN.md.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/idencomp/src unless stated).  The rANS arithmetic lives in the un-vendored crate
 * `rans 0.2.1` -> `ryg-rans-sys 1.0.7` (ryg_rans `rans_byte.h`, public domain, Fabian Giesen 2014); its
 * published algorithm is restated here and pinned by samples/1M.idn.
 */
#ifndef IDN_ORACLE_H
#define IDN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_SCALE_BITS 14 /* idn/model_provider.rs:407 */
#define ORC_ACID_SYMS 5   /* sequence.rs:415 */
#define ORC_Q_SYMS 94     /* fastq/consts.rs:20 */

enum { ORC_TYPE_ACID = 0, ORC_TYPE_QSCORE = 1 };   /* model.rs ModelType repr(u8) */
enum { ORC_KIND_GENERIC = 0, ORC_KIND_LIGHT = 1 }; /* context_spec.rs:218,421; dummy == generic<0,0,0> */

/* error codes (mirror IdnCompressorError / IdnDecompressorError variants) */
enum {
    ORC_OK = 0,
    ORC_E_INVALID_STATE = 1,
    ORC_E_IO = 2,
    ORC_E_SERIALIZE = 3,
    ORC_E_SEQUENCE_TOO_LONG = 4,
    ORC_E_INVALID_VERSION = 5,
    ORC_E_CHECKSUM = 6,
    ORC_E_INVALID_MODEL_INDEX = 7,
    ORC_E_NO_ACTIVE_MODEL = 8,
    ORC_E_UNKNOWN_MODEL = 9,
    ORC_E_UNSUPPORTED = 10,
    ORC_E_FASTQ = 11
};

typedef struct {
    int kind, ao, qo, pb, qmax;
} orc_spec_t;

typedef struct {
    orc_spec_t spec;
    uint32_t base_a, base_q;   /* 5/94 generic, 4/qmax light */
    uint32_t abits, qbits;     /* int_queue.rs:40-43 */
    uint32_t last_pow_a, last_pow_q; /* int_queue.rs:52-59 */
    uint32_t astate, qstate;   /* initial state always 0: int_queue.rs:24-30 */
    uint32_t position, length;
} orc_gen_t;

typedef struct orc_model orc_model_t;

/* ---- a1: quantiser (context.rs:346-394) ---- */
void orc_quantise(const float *probs, int nsym, int scale_bits, uint32_t *cum_out);

/* ---- a4/a5: context-spec generators (context_spec.rs:218-529, int_queue.rs) ---- */
int orc_spec_parse(const char *name, orc_spec_t *out); /* idencomp-macros/src/lib.rs:166-196 */
uint32_t orc_spec_bits(const orc_spec_t *s);
uint64_t orc_spec_num(const orc_spec_t *s);
void orc_gen_init(orc_gen_t *g, const orc_spec_t *s, uint32_t length);
uint32_t orc_gen_current(const orc_gen_t *g);
void orc_gen_update(orc_gen_t *g, uint8_t acid, uint8_t q);

/* ---- a2/a3: model tables (compressor.rs:21-35,109-135; sequence_compressor.rs:21-48,175-201) ---- */
orc_model_t *orc_model_new(int type, const char *spec_name, uint32_t n_ctx, const float *probs,
                           const uint32_t *spec_keys, const uint32_t *spec_ctx, size_t n_specs,
                           const uint8_t identifier[32]);
void orc_model_free(orc_model_t *m);
int orc_model_type(const orc_model_t *m);
uint32_t orc_model_nctx(const orc_model_t *m);
uint32_t orc_model_nsym(const orc_model_t *m);
/* cum row of context index `ctx` (0 = dummy, i+1 = model context i): nsym+1 entries, last = 1<<14 */
void orc_model_cum_row(const orc_model_t *m, uint32_t ctx, uint32_t *out);
uint32_t orc_model_ctx_for(const orc_model_t *m, uint32_t spec);
void orc_model_identifier(const orc_model_t *m, uint8_t out[32]);

/* ---- a8/a9/a6: per-read codec ---- */
/* returns payload length; payload is written to out[0..ret) ; out capacity must be >= 4*len+8 */
size_t orc_encode_read(const orc_model_t *am, const orc_model_t *qm, const uint8_t *acids,
                       const uint8_t *quals, uint32_t len, uint8_t *out);
/* returns 0 on success; final_states[0]=q state, [1]=acid state; consumed = bytes read */
int orc_decode_read(const orc_model_t *am, const orc_model_t *qm, const uint8_t *data, size_t n,
                    uint32_t len, uint8_t *acids, uint8_t *quals, uint32_t final_states[2],
                    size_t *consumed);
size_t orc_score_read(const orc_model_t *m, const uint8_t *acids, const uint8_t *quals, uint32_t len);

/* ---- raw rANS KATs (compressor.rs:224-321) ---- */
/* encode symbols[n] with per-symbol (start,freq) at scale_bits using N interleaved states
 * (symbol i goes to state i % nstates, pushed in the order given); out cap >= 2*n+4*nstates */
size_t orc_rans_encode_raw(const uint32_t *starts, const uint32_t *freqs, size_t n, int nstates,
                           int scale_bits, uint8_t *out);

/* ---- a7/a10/a11: block + container ---- */
typedef struct {
    const orc_model_t *const *models; /* provider order; index = SwitchModel index */
    uint32_t n_models;
    uint32_t max_block_total_len;     /* idn/compressor.rs:187, default 4 Mi */
    int include_identifiers;          /* default 1 */
    int quality;                      /* 1..9, default 7 */
    int fast;                         /* idn/compressor.rs:244-251 */
    int threads;                      /* block worker threads (0/1 = caller thread) */
    int deflate_level;                /* names slice zlib level (6 = flate2 default) */
} orc_params_t;

typedef struct {
    uint64_t n_reads;
    uint64_t total_len;
    const uint64_t *read_off; /* [n_reads+1] offsets into acids/quals */
    const uint8_t *acids;     /* 0..4 */
    const uint8_t *quals;     /* 0..93 */
    const uint64_t *name_off; /* [n_reads+1] offsets into names, may be NULL (= empty names) */
    const uint8_t *names;
} orc_reads_t;

typedef struct {
    uint8_t *data;
    size_t len, cap;
} orc_buf_t;
void orc_buf_free(orc_buf_t *b);

typedef struct {
    uint64_t n_blocks, acid_switches, q_switches, out_acid_bytes, out_q_score_bytes, names_bytes;
} orc_stats_t;

/* whole-file compress (idn/compressor.rs:517-585 + compressor_initializer.rs + compressor_block.rs) */
int orc_compress(const orc_params_t *p, const orc_reads_t *in, orc_buf_t *out, orc_stats_t *stats);

/* one block's slice bytes (no block header), crc returned: compressor_block.rs:83-120 */
int orc_compress_block(const orc_params_t *p, const orc_reads_t *in, uint64_t first_read,
                       uint64_t n_reads, orc_buf_t *out, uint32_t *crc, orc_stats_t *stats);

typedef struct {
    uint64_t n_reads, total_len;
    uint64_t *read_off;
    uint8_t *acids, *quals;
    uint64_t *name_off;
    uint8_t *names;
    uint8_t version;
    uint32_t n_models;
    uint8_t model_ids[256][32];
    uint64_t n_blocks;
} orc_decoded_t;
void orc_decoded_free(orc_decoded_t *d);

/* whole-file decompress (idn/decompressor.rs:304-428 + decompressor_block.rs:77-239).
 * `models` is the caller's provider; it is filtered/reordered by the file's metadata ids. */
int orc_decompress(const orc_model_t *const *models, uint32_t n_models, const uint8_t *idn,
                   size_t idn_len, int threads, orc_decoded_t *out);

/* FASTQ text <-> symbols (fastq/reader.rs:166-282, fastq/writer.rs:190-240, fastq/consts.rs) */
int orc_fastq_parse(const uint8_t *text, size_t n, orc_decoded_t *out);
int orc_fastq_write(const orc_reads_t *in, orc_buf_t *out);

/* GPU-native multi-lane block format (container version 2).  Not part of the reference:
 * an independent CPU statement of our own format, used to check the CUDA implementation of it. */
int orc_compress_native_block(const orc_params_t *p, const orc_reads_t *in, uint64_t first_read, uint64_t n_reads,
                              uint32_t lane_syms, orc_buf_t *out, uint32_t *crc_out);
int orc_native_block_counts(const uint8_t *data, size_t n, uint64_t *n_reads, uint64_t *n_syms);
int orc_decompress_native_block(const orc_model_t *const *models, uint32_t n_models, const uint8_t *data, size_t n,
                                uint8_t *acids, uint8_t *quals, uint32_t *read_len);

/* workload generator (bench/test utility): same sampler as the device library's, see idn_oracle.c */
int orc_synth_reads(const orc_model_t *am, const orc_model_t *qm, const uint64_t *read_off, uint64_t n_reads,
                    uint64_t first_read_index, uint64_t seed, uint32_t n_ppm, int threads, uint8_t *acids,
                    uint8_t *quals);

uint32_t orc_crc32(uint32_t crc, const uint8_t *p, size_t n);
const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
