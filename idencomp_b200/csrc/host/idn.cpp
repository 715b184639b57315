// idn.cpp -- see idn.hpp.  Host orchestration only: every symbol goes through the C-ABI of libidn_gpu.so.
#include "idn.hpp"

#include "clustering.hpp"

#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <future>
#include <numeric>

namespace idencomp {

namespace {

constexpr char kMagic[8] = {'I', 'D', 'E', 'N', 'C', 'O', 'M', 'P'};  // idn/data.rs:3-8

void put_u32be(std::vector<uint8_t>& b, uint32_t v) {
    b.push_back((uint8_t)(v >> 24));
    b.push_back((uint8_t)(v >> 16));
    b.push_back((uint8_t)(v >> 8));
    b.push_back((uint8_t)v);
}
uint32_t get_u32be(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// names slice: raw Deflate at flate2's default level (compressor_block.rs:146-206).  zlib and miniz_oxide emit
// different but equally valid streams, so this slice is compared after inflation.
std::vector<uint8_t> deflate_raw(const uint8_t* src, size_t n) {
    z_stream z;
    std::memset(&z, 0, sizeof z);
    if (deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw IdnError(IDN_E_IO, "deflateInit2 failed");
    std::vector<uint8_t> out(deflateBound(&z, (uLong)n) + 64);
    z.next_in = const_cast<Bytef*>(src);
    z.avail_in = (uInt)n;
    z.next_out = out.data();
    z.avail_out = (uInt)out.size();
    int rc = deflate(&z, Z_FINISH);
    deflateEnd(&z);
    if (rc != Z_STREAM_END) throw IdnError(IDN_E_IO, "deflate failed");
    out.resize(out.size() - z.avail_out);
    return out;
}

std::vector<uint8_t> inflate_raw(const uint8_t* src, size_t n) {
    z_stream z;
    std::memset(&z, 0, sizeof z);
    if (inflateInit2(&z, -15) != Z_OK) throw IdnError(IDN_E_IO, "inflateInit2 failed");
    std::vector<uint8_t> out;
    z.next_in = const_cast<Bytef*>(src);
    z.avail_in = (uInt)n;
    int rc;
    do {
        size_t old = out.size();
        out.resize(old + 65536);
        z.next_out = out.data() + old;
        z.avail_out = 65536;
        rc = inflate(&z, Z_NO_FLUSH);
        out.resize(old + 65536 - z.avail_out);
    } while (rc == Z_OK);
    inflateEnd(&z);
    if (rc != Z_STREAM_END) throw IdnError(IDN_E_SERIALIZE, "identifiers slice does not inflate");
    return out;
}

// Brotli for the identifiers of quality 8-9 (compressor_block.rs:160-170: buffer 4096, quality 11, lgwin 20).  The image
// ships libbrotlienc/libbrotlidec without headers, so the four entry points are bound at run time; when the libraries
// are absent the quality 8-9 identifier path fails with Unsupported.  The byte stream of the C encoder differs from the
// Rust crate's (as zlib's does from miniz_oxide's): this slice is compared after decompression.
struct Brotli {
    using EncFn = int (*)(int quality, int lgwin, int mode, size_t in_size, const uint8_t* in, size_t* out_size, uint8_t* out);
    using BoundFn = size_t (*)(size_t);
    using CreateFn = void* (*)(void*, void*, void*);
    using StreamFn = int (*)(void* state, size_t* avail_in, const uint8_t** next_in, size_t* avail_out, uint8_t** next_out, size_t* total_out);
    using DestroyFn = void (*)(void*);
    EncFn enc = nullptr;
    BoundFn bound = nullptr;
    CreateFn create = nullptr;
    StreamFn stream = nullptr;
    DestroyFn destroy = nullptr;
    Brotli() {
        void* e = dlopen("libbrotlienc.so.1", RTLD_NOW | RTLD_LOCAL);
        void* d = dlopen("libbrotlidec.so.1", RTLD_NOW | RTLD_LOCAL);
        if (e) {
            enc = (EncFn)dlsym(e, "BrotliEncoderCompress");
            bound = (BoundFn)dlsym(e, "BrotliEncoderMaxCompressedSize");
        }
        if (d) {
            create = (CreateFn)dlsym(d, "BrotliDecoderCreateInstance");
            stream = (StreamFn)dlsym(d, "BrotliDecoderDecompressStream");
            destroy = (DestroyFn)dlsym(d, "BrotliDecoderDestroyInstance");
        }
    }
    static Brotli& get() {
        static Brotli b;
        return b;
    }
    std::vector<uint8_t> compress(const uint8_t* src, size_t n) const {
        if (!enc || !bound) throw IdnError(IDN_E_UNSUPPORTED, "Brotli identifier slices (quality 8-9) need libbrotlienc.so.1");
        size_t cap = bound(n) + 64;
        std::vector<uint8_t> out(cap);
        if (!enc(11, 20, 0, n, src, &cap, out.data())) throw IdnError(IDN_E_IO, "Brotli compression failed");
        out.resize(cap);
        return out;
    }
    std::vector<uint8_t> decompress(const uint8_t* src, size_t n) const {
        if (!create || !stream || !destroy) throw IdnError(IDN_E_UNSUPPORTED, "Brotli identifier slices need libbrotlidec.so.1");
        void* st = create(nullptr, nullptr, nullptr);
        if (!st) throw IdnError(IDN_E_IO, "BrotliDecoderCreateInstance failed");
        std::vector<uint8_t> out(std::max<size_t>(65536, 8 * n));
        size_t avail_in = n, total = 0;
        const uint8_t* next_in = src;
        int rc;
        for (;;) {
            size_t avail_out = out.size() - total;
            uint8_t* next_out = out.data() + total;
            rc = stream(st, &avail_in, &next_in, &avail_out, &next_out, nullptr);
            total = out.size() - avail_out;
            if (rc != 3) break;  // BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT
            out.resize(out.size() * 2);
        }
        destroy(st);
        if (rc != 1) throw IdnError(IDN_E_SERIALIZE, "identifiers slice is not a valid Brotli stream");  // 1 = SUCCESS
        out.resize(total);
        return out;
    }
};

}  // namespace

// ---- DeviceModels ---------------------------------------------------------------------------------------------------
DeviceModels::~DeviceModels() {
    if (ctx_) idn_gpu_destroy(ctx_);
}
void DeviceModels::open(int32_t device) {
    if (ctx_) return;
    int32_t rc = idn_gpu_create(device, &ctx_);
    if (rc) throw IdnError(rc, "no usable CUDA device (this implementation has no CPU fallback)");
}
void DeviceModels::raise(int32_t rc) const { throw IdnError(rc, ctx_ ? idn_gpu_last_error(ctx_) : "no device context"); }
void DeviceModels::upload(const ModelProvider& provider) {
    for (idn_model_t h : handles_) idn_gpu_model_release(ctx_, h);
    handles_.clear();
    for (size_t i = 0; i < provider.len(); i++) {
        idn_model_t h = -1;
        int32_t rc = upload_model(ctx_, provider[i], &h);
        if (rc) raise(rc);
        handles_.push_back(h);
    }
}

IdnCompressorParamsBuilder& IdnCompressorParamsBuilder::quality(uint8_t v) {
    if (v < 1 || v > 9) throw std::invalid_argument("compression quality must be between 1 and 9");
    p_.quality = v;
    return *this;
}

// ---- IdnCompressor ---------------------------------------------------------------------------------------------------
IdnCompressor::IdnCompressor(Sink sink, IdnCompressorParams params) : sink_(std::move(sink)), params_(std::move(params)) {
    if (params_.fast) params_.quality = 1;  // IdnCompressorParamsBuilder::fast (idn/compressor.rs:244-251)
    if (params_.mode != IDN_MODE_COMPAT && params_.mode != IDN_MODE_NATIVE) throw IdnError(IDN_E_INVALID_STATE, "unknown container mode");
    dev_.open(params_.device);
    if (params_.mode == IDN_MODE_NATIVE) {
        int32_t rc = idn_gpu_set_lane_symbols(dev_.ctx(), params_.lane_symbols);
        if (rc) dev_.raise(rc);
    }
}

IdnCompressor::~IdnCompressor() = default;

void IdnCompressor::add_sequence(FastqSequence seq) {
    if (finished_) throw IdnError(IDN_E_INVALID_STATE, "add_sequence after finish");
    if (seq.acids.size() != seq.quality_scores.size()) throw IdnError(IDN_E_INVALID_STATE, "acids and quality scores differ in length");
    const uint64_t len = seq.len();
    if (len > params_.max_block_total_len / 2)  // idn/compressor.rs:542-544
        throw IdnError(IDN_E_SEQUENCE_TOO_LONG, "sequence too long: " + std::to_string(len) + " > " + std::to_string(params_.max_block_total_len / 2));
    if (cur_block_len_ + len > params_.max_block_total_len) make_block();  // :526-540
    acids_.insert(acids_.end(), seq.acids.begin(), seq.acids.end());
    quals_.insert(quals_.end(), seq.quality_scores.begin(), seq.quality_scores.end());
    read_off_.push_back(acids_.size());
    if (params_.include_identifiers) names_.insert(names_.end(), seq.identifier.begin(), seq.identifier.end());
    name_off_.push_back(names_.size());
    cur_block_len_ += len;
    stats_.in_symbols += len;
    stats_.in_reads++;
    stats_.in_identifier_bytes += seq.identifier.size();
}

void IdnCompressor::make_block() {
    const uint32_t n_reads = (uint32_t)(read_off_.size() - 1);
    if (n_reads == block_first_.back()) return;  // nothing since the last block
    block_first_.push_back(n_reads);
    cur_block_len_ = 0;
    if (block_first_.size() - 1 >= params_.batch_blocks) flush_batch();
}

std::vector<ModelIdentifier> IdnCompressor::best_models(ModelType type, size_t model_num, const std::vector<uint32_t>& sizes,
                                                        const std::vector<size_t>& cols, size_t n_cols, size_t n_reads,
                                                        Clustering& clustering) {
    // cost matrix of this type's models only
    const size_t n = cols.size();
    std::vector<uint32_t> cost(n_reads * n);
    for (size_t r = 0; r < n_reads; r++)
        for (size_t k = 0; k < n; k++) cost[r * n + k] = sizes[r * n_cols + cols[k]];
    std::vector<size_t> pick = params_.quality >= 2 ? clustering.make_clusters(cost, n_reads, n, model_num, nullptr)  // CLUSTERING_THRESHOLD
                                                    : rank_models(cost, n_reads, n, model_num);
    std::vector<ModelIdentifier> ids;
    for (size_t k : pick) ids.push_back(params_.model_provider[cols[k]].identifier());
    (void)type;
    return ids;
}

void IdnCompressor::initialize() {
    ModelProvider& mp = params_.model_provider;
    std::vector<size_t> of_type[2];
    for (size_t i = 0; i < mp.len(); i++) of_type[(size_t)mp[i].model_type()].push_back(i);
    if (of_type[0].empty() || of_type[1].empty()) throw IdnError(IDN_E_INVALID_STATE, "the model provider needs at least one model per type");
    const size_t model_num = ((size_t)params_.quality + 1) / 2;  // compressor_initializer.rs:56
    std::vector<ModelIdentifier> ids;
    if (of_type[0].size() == 1 && of_type[1].size() == 1) {  // "Only one model registered" (model_chooser.rs:37-40)
        ids = {mp[of_type[0][0]].identifier(), mp[of_type[1][0]].identifier()};
    } else {
        // cost matrix over the reads of the FIRST block with every model of the provider (a6 on the device)
        dev_.upload(mp);
        const uint64_t n_reads = block_first_.size() > 1 ? block_first_[1] : read_off_.size() - 1;
        std::vector<uint32_t> sizes(n_reads * mp.len());
        idn_batch b{};
        b.n_reads = n_reads;
        b.n_symbols = read_off_[n_reads];
        b.acids = acids_.data();
        b.quals = quals_.data();
        b.read_off = read_off_.data();
        if (n_reads) {
            int32_t rc = idn_gpu_score(dev_.ctx(), &b, dev_.handles().data(), (uint32_t)mp.len(), sizes.data());
            if (rc) dev_.raise(rc);
        }
        // ONE Clustering (one random stream) serves the acid models and then the q-score models, as the reference's
        // ModelChooser does (model_chooser.rs:14-24, compressor_initializer.rs:57-64)
        Clustering clustering;
        for (int t = 0; t < 2; t++) {
            std::vector<ModelIdentifier> got;
            if (of_type[t].size() == 1) got = {mp[of_type[t][0]].identifier()};
            else got = best_models((ModelType)t, model_num, sizes, of_type[t], mp.len(), n_reads, clustering);
            ids.insert(ids.end(), got.begin(), got.end());
        }
    }
    mp.filter_by_identifiers(ids);  // acid ids first, then q-score ids (compressor_initializer.rs:57-74)
    if (mp.len() > IDN_MAX_MODELS) throw IdnError(IDN_E_UNSUPPORTED, "too many retained models");
    dev_.upload(mp);
    retained_ = ids;
    // header + metadata (writer_idn.rs:25-59, data.rs:3-33)
    std::vector<uint8_t> h(kMagic, kMagic + 8);
    h.push_back((uint8_t)(params_.mode == IDN_MODE_NATIVE ? 2 : 1));
    h.push_back(1);  // item_num
    h.push_back(0);  // IdnMetadataItem::Models
    h.push_back((uint8_t)ids.size());
    for (auto& id : ids) h.insert(h.end(), id.begin(), id.end());
    sink_(h.data(), h.size());
    stats_.out_bytes += h.size();
    initialized_ = true;
}

void IdnCompressor::flush_batch() {
    const uint32_t n_blocks = (uint32_t)(block_first_.size() - 1);
    if (n_blocks == 0) return;
    if (!initialized_) initialize();
    const uint64_t n_reads = block_first_.back();
    // identifiers slices on host threads, like write_identifiers (compressor_block.rs:146-206): names joined by '\n'
    std::vector<std::vector<uint8_t>> name_slices(n_blocks);
    std::vector<uint32_t> prefix(n_blocks, 0);
    if (params_.include_identifiers) {
        const bool brotli = params_.quality >= 8;  // BROTLI_THRESHOLD (compressor_block.rs:146)
        if (brotli && !Brotli::get().enc)  // fail early, on the caller's thread, when the library is missing
            throw IdnError(IDN_E_UNSUPPORTED, "Brotli identifier slices (quality 8-9) need libbrotlienc.so.1");
        auto one = [&](uint32_t b) {
            std::vector<uint8_t> joined;
            for (uint32_t r = block_first_[b]; r < block_first_[b + 1]; r++) {
                if (r > block_first_[b]) joined.push_back('\n');
                joined.insert(joined.end(), names_.begin() + name_off_[r], names_.begin() + name_off_[r + 1]);
            }
            std::vector<uint8_t> z = brotli ? Brotli::get().compress(joined.data(), joined.size()) : deflate_raw(joined.data(), joined.size());
            std::vector<uint8_t> s;
            s.push_back(0x00);  // IdnSliceHeader::Identifiers (data.rs:46-47)
            put_u32be(s, (uint32_t)z.size());
            s.push_back(brotli ? 0 : 1);  // IdnIdentifierCompression::{Brotli = 0, Deflate = 1} (data.rs:57-61)
            s.insert(s.end(), z.begin(), z.end());
            name_slices[b] = std::move(s);
        };
        if (params_.thread_num > 1) {
            std::vector<std::future<void>> jobs;
            for (uint32_t b = 0; b < n_blocks; b++) jobs.push_back(std::async(std::launch::async, one, b));
            for (auto& j : jobs) j.get();
        } else {
            for (uint32_t b = 0; b < n_blocks; b++) one(b);
        }
        for (uint32_t b = 0; b < n_blocks; b++) prefix[b] = (uint32_t)name_slices[b].size();
    }
    uint64_t prefix_total = std::accumulate(prefix.begin(), prefix.end(), (uint64_t)0);
    idn_batch b{};
    b.n_reads = n_reads;
    b.n_symbols = read_off_[n_reads];
    b.n_blocks = n_blocks;
    b.acids = acids_.data();
    b.quals = quals_.data();
    b.read_off = read_off_.data();
    b.block_first_read = block_first_.data();
    if (params_.include_identifiers) {
        b.names = names_.empty() ? reinterpret_cast<const uint8_t*>("") : names_.data();
        b.name_off = name_off_.data();
    }
    out_.resize(idn_gpu_compress_bound(n_reads, b.n_symbols, n_blocks, prefix_total));
    std::vector<uint64_t> block_off(n_blocks + 1);
    idn_compress_stats st{};
    int32_t rc = idn_gpu_compress_blocks(dev_.ctx(), &b, params_.mode, dev_.handles().data(), (uint32_t)dev_.handles().size(),
                                         params_.fast ? 1 : 0, params_.include_identifiers ? prefix.data() : nullptr, out_.data(),
                                         out_.size(), block_off.data(), nullptr, &st);
    if (rc) dev_.raise(rc);
    for (uint32_t k = 0; k < n_blocks; k++)
        if (prefix[k]) std::memcpy(out_.data() + block_off[k] + 8, name_slices[k].data(), prefix[k]);
    sink_(out_.data(), st.out_bytes);  // blocks in order: this replaces IdnBlockLock (common.rs:10-57)
    stats_.out_bytes += st.out_bytes;
    stats_.out_identifier_bytes += prefix_total;
    stats_.out_payload_bytes += st.payload_bytes;
    stats_.acid_model_switches += st.acid_switches;
    stats_.q_score_model_switches += st.q_switches;
    stats_.blocks += n_blocks;
    // keep the reads of an unfinished block (none: flush happens at block boundaries) and reset the batch
    acids_.clear();
    quals_.clear();
    names_.clear();
    read_off_.assign(1, 0);
    name_off_.assign(1, 0);
    block_first_.assign(1, 0);
}

void IdnCompressor::finish() {
    if (finished_) throw IdnError(IDN_E_INVALID_STATE, "finish called twice");
    make_block();  // flush the partial block (idn/compressor.rs:575-578)
    flush_batch();
    if (!initialized_) initialize();  // empty file: header + metadata still get written
    const uint8_t terminator[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // an empty block marks the end (:579)
    sink_(terminator, 8);
    stats_.out_bytes += 8;
    finished_ = true;
}

// ---- IdnDecompressor -------------------------------------------------------------------------------------------------
IdnDecompressor::IdnDecompressor(Source source, IdnDecompressorParams params) : source_(std::move(source)), params_(std::move(params)) {
    dev_.open(params_.device);
    if (params_.batch_blocks == 0) params_.batch_blocks = 1;
}
IdnDecompressor::~IdnDecompressor() = default;

void IdnDecompressor::read_exact(uint8_t* dst, size_t n, const char* what) {
    size_t got = 0;
    while (got < n) {
        size_t k = source_(dst + got, n - got);
        if (k == 0) throw IdnError(IDN_E_IO, std::string("unexpected end of input while reading ") + what);
        got += k;
    }
}

void IdnDecompressor::initialize() {
    uint8_t h[9];
    read_exact(h, 9, "the header");
    if (std::memcmp(h, kMagic, 8) != 0) throw IdnError(IDN_E_SERIALIZE, "not an IDN file (bad magic)");
    version_ = h[8];
    if (version_ != 1 && version_ != 2)  // idn/decompressor.rs:317-319 (version 2 = this implementation's native format)
        throw IdnError(IDN_E_INVALID_VERSION, "unsupported IDN version " + std::to_string(version_));
    uint8_t item_num;
    read_exact(&item_num, 1, "the metadata");
    std::vector<ModelIdentifier> ids;
    for (uint8_t i = 0; i < item_num; i++) {
        uint8_t kind;
        read_exact(&kind, 1, "a metadata item");
        if (kind != 0) throw IdnError(IDN_E_SERIALIZE, "unknown metadata item");
        uint8_t n;
        read_exact(&n, 1, "the model list");
        for (uint8_t k = 0; k < n; k++) {
            ModelIdentifier id;
            read_exact(id.data(), 32, "a model identifier");
            ids.push_back(id);
        }
    }
    if (!params_.model_provider.has_all_models(ids)) {  // IdnDecompressorError::UnknownModel (idn/decompressor.rs:340-350)
        for (auto& id : ids)
            if (!params_.model_provider.has_all_models({id})) throw IdnError(IDN_E_UNKNOWN_MODEL, "unknown model " + to_hex(id));
    }
    params_.model_provider.filter_by_identifiers(ids);
    dev_.upload(params_.model_provider);
    initialized_ = true;
}

bool IdnDecompressor::read_batch() {
    // block headers + payloads of up to batch_blocks blocks into one buffer (idn/decompressor.rs:387-428)
    std::vector<uint8_t> buf;
    std::vector<uint64_t> off;
    std::vector<uint32_t> len, crc;
    bool more = true;
    while (off.size() < params_.batch_blocks) {
        uint8_t h[8];
        read_exact(h, 8, "a block header");
        uint32_t n = get_u32be(h), c = get_u32be(h + 4);
        if (n == 0) {  // terminator (:422-425)
            more = false;
            break;
        }
        off.push_back(buf.size());
        len.push_back(n);
        crc.push_back(c);
        buf.resize(buf.size() + n);
        read_exact(buf.data() + off.back(), n, "a block");
    }
    const uint32_t n_blocks = (uint32_t)off.size();
    if (n_blocks == 0) return more;
    off.push_back(buf.size());
    const int32_t mode = version_ == 2 ? IDN_MODE_NATIVE : IDN_MODE_COMPAT;
    const auto& handles = dev_.handles();
    idn_block_index_totals tot{};
    std::vector<uint32_t> block_first(n_blocks + 1);
    int32_t rc = idn_gpu_index_blocks(dev_.ctx(), buf.data(), off.data(), len.data(), n_blocks, mode, handles.data(),
                                      (uint32_t)handles.size(), &tot, block_first.data());
    if (rc) dev_.raise(rc);
    // identifiers: leading Identifiers slices of every block, inflated on the host (decompressor_block.rs:146-192)
    std::vector<uint8_t> names;
    std::vector<uint64_t> name_off(tot.n_reads + 1, 0);
    bool any_names = false;
    for (uint32_t b = 0; b < n_blocks; b++) {
        const uint8_t* p = buf.data() + off[b];
        size_t pos = 0;
        uint64_t r = block_first[b];
        const uint64_t r_end = block_first[b + 1];
        while (pos < len[b] && p[pos] == 0x00) {
            if (pos + 6 > len[b]) throw IdnError(IDN_E_SERIALIZE, "truncated identifiers slice");
            uint32_t n = get_u32be(p + pos + 1);
            uint8_t comp = p[pos + 5];
            if (n > len[b] - pos - 6) throw IdnError(IDN_E_SERIALIZE, "truncated identifiers slice");
            if (comp > 1) throw IdnError(IDN_E_SERIALIZE, "unknown identifier compression");
            std::vector<uint8_t> text = comp == 0 ? Brotli::get().decompress(p + pos + 6, n) : inflate_raw(p + pos + 6, n);
            any_names = true;
            // split at '\n': one identifier per sequence, in order (identifiers_as_lines)
            size_t s = 0;
            while (r < r_end) {
                size_t e = s;
                while (e < text.size() && text[e] != '\n') e++;
                names.insert(names.end(), text.begin() + s, text.begin() + e);
                name_off[++r] = names.size();
                if (e >= text.size()) break;
                s = e + 1;
            }
            pos += 6 + (size_t)n;
        }
        for (; r < r_end; r++) name_off[r + 1] = names.size();  // sequences without an identifier
    }
    std::vector<uint8_t> acids(tot.n_symbols + 1), quals(tot.n_symbols + 1);
    std::vector<uint64_t> read_off(tot.n_reads + 1, 0);
    int32_t bad = -1;
    if (names.empty()) names.push_back(0);
    rc = idn_gpu_decompress_blocks(dev_.ctx(), buf.data(), off.data(), len.data(), crc.data(), n_blocks, mode, handles.data(),
                                   (uint32_t)handles.size(), any_names ? names.data() : nullptr, any_names ? name_off.data() : nullptr,
                                   acids.data(), quals.data(), read_off.data(), tot.n_reads, tot.n_symbols, &bad);
    if (rc) dev_.raise(rc);
    for (uint64_t r = 0; r < tot.n_reads; r++) {
        FastqSequence s;
        if (any_names) s.identifier.assign(names.begin() + name_off[r], names.begin() + name_off[r + 1]);
        s.acids.assign(acids.begin() + read_off[r], acids.begin() + read_off[r + 1]);
        s.quality_scores.assign(quals.begin() + read_off[r], quals.begin() + read_off[r + 1]);
        queue_.push_back(std::move(s));
    }
    return more;
}

std::optional<FastqSequence> IdnDecompressor::next_sequence() {
    if (!initialized_) initialize();
    while (queue_.empty() && !eof_) eof_ = !read_batch();
    if (queue_.empty()) return std::nullopt;
    FastqSequence s = std::move(queue_.front());
    queue_.pop_front();
    return s;
}

}  // namespace idencomp
