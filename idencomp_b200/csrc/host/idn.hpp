// idn.hpp -- host-side mirror of idencomp's compressor / decompressor API (the drop-in surface above the C-ABI).
//
//   FastqSequence                 fastq/mod.rs (identifier, acids, quality scores)
//   IdnCompressorParams(Builder)  idn/compressor.rs:164-274   model_provider, max_block_total_len, thread_num,
//                                                             include_identifiers, quality, fast  (+ device, mode, batch_blocks)
//   IdnCompressor                 idn/compressor.rs:443-585   with_params, add_sequence, finish
//   IdnDecompressorParams         idn/decompressor.rs:167-229
//   IdnDecompressor               idn/decompressor.rs:455-566 with_params, next_sequence
//   IdnError                      idn/compressor.rs:22-33, idn/decompressor.rs:25-48 (same variants, same conditions)
//
// Everything below add_sequence / next_sequence that touches symbols runs on the GPU through include/idn_gpu.h; the
// host keeps what the reference keeps on the host: block forming, container framing, name Deflate, ordered writes.
// The reference hands one block to one thread-pool job; here `batch_blocks` blocks go to the device in one call.
#pragma once
#include <cstdint>
#include <deque>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/idn_gpu.h"
#include "model.hpp"

namespace idencomp {

struct FastqSequence {
    std::string identifier;
    std::vector<uint8_t> acids;           // Acid as u8: N=0 A=1 C=2 T=3 G=4 (sequence.rs:401-413)
    std::vector<uint8_t> quality_scores;  // 0..93
    size_t len() const { return acids.size(); }
};

// IdnCompressorError / IdnDecompressorError: `code` is the IDN_* status of include/idn_gpu.h
struct IdnError : std::runtime_error {
    int32_t code;
    IdnError(int32_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

struct CompressionStats {  // idn/compressor.rs:596-737 (the counters that exist on this path)
    uint64_t in_symbols = 0, in_reads = 0, in_identifier_bytes = 0;
    uint64_t out_bytes = 0, out_identifier_bytes = 0, out_payload_bytes = 0;
    uint64_t blocks = 0, acid_model_switches = 0, q_score_model_switches = 0;
};

struct IdnCompressorParams {
    ModelProvider model_provider;
    uint32_t max_block_total_len = 4 * 1024 * 1024;  // idn/compressor.rs:187
    uint32_t thread_num = 0;                          // host threads for the name codec (0 = caller thread)
    bool include_identifiers = true;
    uint8_t quality = 7;                              // 1..9
    bool fast = false;
    // additions of this implementation (defaults keep the reference's behaviour)
    int32_t device = 0;
    int32_t mode = IDN_MODE_COMPAT;  // IDN_MODE_NATIVE writes container version 2
    uint32_t batch_blocks = 32;      // blocks per device call
    uint32_t lane_symbols = 2048;    // native mode lane quantum
};

class IdnCompressorParamsBuilder {
public:
    IdnCompressorParamsBuilder& model_provider(ModelProvider p) { p_.model_provider = std::move(p); return *this; }
    IdnCompressorParamsBuilder& max_block_total_len(uint32_t v) { p_.max_block_total_len = v; return *this; }
    IdnCompressorParamsBuilder& thread_num(uint32_t v) { p_.thread_num = v; return *this; }
    IdnCompressorParamsBuilder& include_identifiers(bool v) { p_.include_identifiers = v; return *this; }
    IdnCompressorParamsBuilder& quality(uint8_t v);  // throws std::invalid_argument outside 1..9 (CompressionQuality::new)
    IdnCompressorParamsBuilder& fast(bool v) { p_.fast = v; return *this; }
    IdnCompressorParamsBuilder& device(int32_t v) { p_.device = v; return *this; }
    IdnCompressorParamsBuilder& mode(int32_t v) { p_.mode = v; return *this; }
    IdnCompressorParamsBuilder& batch_blocks(uint32_t v) { p_.batch_blocks = v ? v : 1; return *this; }
    IdnCompressorParamsBuilder& lane_symbols(uint32_t v) { p_.lane_symbols = v; return *this; }
    IdnCompressorParams build() const { return p_; }

private:
    IdnCompressorParams p_;
};

// RAII over idn_gpu_ctx + the handles of the uploaded models of one provider
class DeviceModels {
public:
    DeviceModels() = default;
    ~DeviceModels();
    DeviceModels(const DeviceModels&) = delete;
    DeviceModels& operator=(const DeviceModels&) = delete;
    void open(int32_t device);
    void upload(const ModelProvider& provider);  // replaces the current set
    idn_gpu_ctx* ctx() const { return ctx_; }
    const std::vector<idn_model_t>& handles() const { return handles_; }
    [[noreturn]] void raise(int32_t rc) const;  // IdnError from idn_gpu_last_error

private:
    idn_gpu_ctx* ctx_ = nullptr;
    std::vector<idn_model_t> handles_;
};

class IdnCompressor {
public:
    using Sink = std::function<void(const uint8_t*, size_t)>;  // the reference's `W: Write`
    IdnCompressor(Sink sink, IdnCompressorParams params);  // IdnCompressor::with_params
    ~IdnCompressor();
    void add_sequence(FastqSequence seq);  // SequenceTooLong when len > max_block_total_len / 2 (idn/compressor.rs:542-544)
    void finish();                         // flushes, writes the empty terminator block; InvalidState if called twice
    const CompressionStats& stats() const { return stats_; }
    // the model identifiers written to the metadata (acid ids first), available after the first block was processed
    const std::vector<ModelIdentifier>& retained_models() const { return retained_; }

private:
    struct Batch;
    void make_block();
    void flush_batch();
    void initialize();  // CompressorInitializer::initialize (idn/compressor_initializer.rs:33-74)
    std::vector<ModelIdentifier> best_models(ModelType type, size_t model_num, const std::vector<uint32_t>& sizes,
                                             const std::vector<size_t>& cols, size_t n_cols, size_t n_reads, class Clustering& clustering);

    Sink sink_;
    IdnCompressorParams params_;
    DeviceModels dev_;
    CompressionStats stats_;
    std::vector<ModelIdentifier> retained_;
    bool initialized_ = false, finished_ = false;
    // SoA batch under construction
    std::vector<uint8_t> acids_, quals_, names_;
    std::vector<uint64_t> read_off_{0}, name_off_{0};
    std::vector<uint32_t> block_first_{0};
    uint64_t cur_block_len_ = 0;
    std::vector<uint8_t> out_;
};

struct IdnDecompressorParams {
    ModelProvider model_provider;
    uint32_t thread_num = 0;
    int32_t device = 0;
    uint32_t batch_blocks = 32;
};

class IdnDecompressor {
public:
    using Source = std::function<size_t(uint8_t*, size_t)>;  // reads up to n bytes, returns the count (0 = end of input)
    IdnDecompressor(Source source, IdnDecompressorParams params);  // IdnDecompressor::with_params
    ~IdnDecompressor();
    std::optional<FastqSequence> next_sequence();  // None at the end of the file
    uint8_t version() const { return version_; }

private:
    void initialize();  // header + metadata (idn/decompressor.rs:304-374)
    bool read_batch();  // false once the terminator block was seen
    void read_exact(uint8_t* dst, size_t n, const char* what);

    Source source_;
    IdnDecompressorParams params_;
    DeviceModels dev_;
    bool initialized_ = false, eof_ = false;
    uint8_t version_ = 0;
    std::deque<FastqSequence> queue_;
};

}  // namespace idencomp
