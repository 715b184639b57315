// model.cpp -- see model.hpp.  Compile WITHOUT -ffast-math / FMA contraction: quantise() must round like the reference.
#include "model.hpp"

#include <dirent.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "msgpack_lite.hpp"
#include "sha3.hpp"

namespace idencomp {

// ---- ContextSpecType ------------------------------------------------------------------------------------------------
static bool parse_int(const char*& p, int* out) {
    if (*p < '0' || *p > '9') return false;
    long v = 0;
    while (*p >= '0' && *p <= '9') {
        v = v * 10 + (*p - '0');
        if (v > 1000) return false;
        p++;
    }
    *out = (int)v;
    return true;
}
static bool eat(const char*& p, const char* lit) {
    size_t n = std::strlen(lit);
    if (std::strncmp(p, lit, n) != 0) return false;
    p += n;
    return true;
}

ContextSpecType ContextSpecType::parse(const std::string& name) {
    ContextSpecType s;
    s.name = name;
    if (name == "dummy") return s;
    const char* p = name.c_str();
    bool ok = false;
    if (eat(p, "generic_ao")) {
        s.kind = Generic;
        ok = parse_int(p, &s.acid_order) && eat(p, "_qo") && parse_int(p, &s.q_score_order) && eat(p, "_pb") &&
             parse_int(p, &s.position_bits) && *p == 0;
    } else if (eat(p, "light_ao")) {
        s.kind = Light;
        ok = parse_int(p, &s.acid_order) && eat(p, "_qo") && parse_int(p, &s.q_score_order) && eat(p, "_pb") &&
             parse_int(p, &s.position_bits) && eat(p, "_qm") && parse_int(p, &s.q_score_max) && *p == 0 &&
             s.q_score_max >= 1 && s.q_score_max <= 94;
    }
    if (!ok || s.acid_order > 8 || s.q_score_order > 8 || s.position_bits > 16)
        throw ModelError("unknown context spec type `" + name + "`");
    if (s.bits() > 31) throw ModelError("context spec type `" + name + "` does not fit 32 bits");
    return s;
}

static uint32_t queue_bits(uint64_t base, int order) {  // IntQueue::num_bits, int_queue.rs:40-43
    uint64_t v = 1;
    for (int i = 0; i < order; i++) v *= base;
    v -= 1;
    uint32_t n = 0;
    while (v) {
        n++;
        v >>= 1;
    }
    return n;
}

uint32_t ContextSpecType::bits() const {
    uint64_t ba = kind == Light ? 4 : 5, bq = kind == Light ? (uint64_t)q_score_max : 94;
    return queue_bits(ba, acid_order) + queue_bits(bq, q_score_order) + (uint32_t)position_bits;
}
uint64_t ContextSpecType::spec_num() const { return 1ull << bits(); }

std::string to_hex(const ModelIdentifier& id) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (uint8_t b : id) {
        s.push_back(d[b >> 4]);
        s.push_back(d[b & 15]);
    }
    return s;
}

// ---- quantiser -------------------------------------------------------------------------------------------------------
std::vector<uint32_t> quantise(const float* probs, size_t nsym, uint32_t scale_bits) {
    const uint32_t total = 1u << scale_bits;
    if (total <= nsym) throw ModelError("scale_bits too small for the alphabet");
    std::vector<uint32_t> v(nsym);
    volatile float acc = 0.0f;  // volatile: every partial sum is rounded to f32, as the reference's scan does
    for (size_t i = 0; i < nsym; i++) {
        float raw = acc;
        volatile float term = probs[i] * (float)total;
        acc = acc + term;
        v[i] = (uint32_t)std::roundf(raw);  // f32::round: half away from zero
    }
    // cum -> freq (context.rs:411-417)
    for (size_t i = 0; i + 1 < nsym; i++) v[i] = v[i + 1] - v[i];
    v[nsym - 1] = total - v[nsym - 1];
    // fix_zero_freqs (context.rs:373-394)
    uint32_t zero_count = 0;
    for (auto& f : v)
        if (f == 0) {
            f = 1;
            zero_count++;
        }
    size_t i = 0;
    while (zero_count > 0) {
        if (v[i] > 1) {
            v[i]--;
            zero_count--;
        }
        if (++i >= nsym) i = 0;
    }
    // freq -> cum
    uint32_t run = 0;
    for (auto& f : v) {
        uint32_t t = f;
        f = run;
        run += t;
    }
    if (run != total) throw ModelError("quantised frequencies do not sum to the total");
    return v;
}

// ---- Model -----------------------------------------------------------------------------------------------------------
Model::Model(ModelType type, ContextSpecType spec_type, std::vector<ModelContext> contexts)
    : type_(type), spec_(std::move(spec_type)), contexts_(std::move(contexts)) {
    const size_t nsym = symbols_of(type_);
    for (auto& c : contexts_) {
        if (c.symbol_prob.size() != nsym) throw ModelError("context has the wrong number of symbols");
        std::sort(c.specs.begin(), c.specs.end());
    }
    if (contexts_.size() > 65536) throw ModelError("model has more than 65536 contexts");  // check_model, sequence_compressor.rs:209-219
    std::stable_sort(contexts_.begin(), contexts_.end(),
                     [](const ModelContext& a, const ModelContext& b) { return a.specs < b.specs; });  // map_contexts
    const uint64_t n = spec_.spec_num();
    for (auto& c : contexts_)
        for (uint32_t s : c.specs)
            if (s >= n) throw ModelError("context spec out of range for the spec type");
    make_identifier();
}

Model Model::empty(ModelType type) { return Model(type, ContextSpecType::parse("dummy"), {}); }

static void put_be32(std::vector<uint8_t>& b, uint32_t v) {
    b.push_back((uint8_t)(v >> 24));
    b.push_back((uint8_t)(v >> 16));
    b.push_back((uint8_t)(v >> 8));
    b.push_back((uint8_t)v);
}

void Model::make_identifier() {
    Sha3_256 h;
    uint8_t t = (uint8_t)type_;
    h.update(&t, 1);
    h.update(spec_.name.data(), spec_.name.size());
    std::vector<uint8_t> buf;
    for (auto& c : contexts_)
        for (float p : c.symbol_prob) {
            uint32_t u;
            std::memcpy(&u, &p, 4);
            put_be32(buf, u);
        }
    h.update(buf.data(), buf.size());
    std::vector<std::pair<uint32_t, uint32_t>> entries;
    for (size_t i = 0; i < contexts_.size(); i++)
        for (uint32_t s : contexts_[i].specs) entries.emplace_back(s, (uint32_t)i);
    std::sort(entries.begin(), entries.end());
    buf.clear();
    for (auto& e : entries) {
        put_be32(buf, e.first);
        put_be32(buf, e.second);
    }
    h.update(buf.data(), buf.size());
    id_ = h.finish();
}

Model Model::read_msgpack(const uint8_t* data, size_t n) {
    try {
        MsgpackReader r(data, n);
        if (r.read_array() != 4) throw ModelError("model file: expected a 4-field record");
        ModelIdentifier stored;
        r.read_bytes(stored.data(), 32);
        std::string type_s = r.read_str();
        ModelType type;
        if (type_s == "Acids") type = ModelType::Acids;
        else if (type_s == "QualityScores") type = ModelType::QualityScores;
        else throw ModelError("model file: unknown model type `" + type_s + "`");
        ContextSpecType spec = ContextSpecType::parse(r.read_str());
        size_t n_ctx = r.read_array();
        std::vector<ModelContext> ctxs(n_ctx);
        for (auto& c : ctxs) {
            if (r.read_array() != 2) throw ModelError("model file: bad context record");
            size_t ns = r.read_array();
            c.specs.resize(ns);
            for (auto& s : c.specs) s = (uint32_t)r.read_uint();
            if (r.read_array() != 2) throw ModelError("model file: bad context body");
            c.context_prob = r.read_f32();
            size_t np = r.read_array();
            c.symbol_prob.resize(np);
            for (auto& p : c.symbol_prob) p = r.read_f32();
        }
        Model m(type, spec, std::move(ctxs));
        if (m.identifier() != stored)  // model_serializer.rs:111-114
            throw ModelError("model file: stored identifier " + to_hex(stored) + " != computed " + to_hex(m.identifier()));
        return m;
    } catch (const MsgpackError& e) {
        throw ModelError(std::string("model file: ") + e.what());
    }
}

Model Model::read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw ModelError("cannot open `" + path + "`");
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return read_msgpack(buf.data(), buf.size());
}

std::vector<uint16_t> Model::cum_table() const {
    const size_t nsym = symbols_of(type_);
    std::vector<uint16_t> out((contexts_.size() + 1) * (nsym + 1));
    auto put_row = [&](size_t row, const float* probs) {
        std::vector<uint32_t> cum = quantise(probs, nsym, kScaleBits);
        uint16_t* o = out.data() + row * (nsym + 1);
        for (size_t i = 0; i < nsym; i++) o[i] = (uint16_t)cum[i];
        o[nsym] = (uint16_t)(1u << kScaleBits);
    };
    std::vector<float> dummy(nsym, 1.0f / (float)nsym);  // Context::dummy, context.rs:225-229
    put_row(0, dummy.data());
    for (size_t i = 0; i < contexts_.size(); i++) put_row(i + 1, contexts_[i].symbol_prob.data());
    return out;
}

void Model::spec_table(std::vector<uint32_t>* keys, std::vector<uint32_t>* ctx) const {
    keys->clear();
    ctx->clear();
    for (size_t i = 0; i < contexts_.size(); i++)
        for (uint32_t s : contexts_[i].specs) {
            keys->push_back(s);
            ctx->push_back((uint32_t)i);
        }
}

int32_t upload_model(idn_gpu_ctx* ctx, const Model& m, idn_model_t* handle) {
    const ContextSpecType& st = m.context_spec_type();
    std::vector<uint16_t> cum = m.cum_table();
    std::vector<uint32_t> keys, cidx;
    m.spec_table(&keys, &cidx);
    return idn_gpu_model_upload(ctx, (int32_t)m.model_type(), st.kind == ContextSpecType::Light ? IDN_SPEC_LIGHT : IDN_SPEC_GENERIC,
                                st.acid_order, st.q_score_order, st.position_bits, st.q_score_max, (uint32_t)m.len(), cum.data(),
                                keys.data(), cidx.data(), keys.size(), handle);
}

// ---- ModelProvider ---------------------------------------------------------------------------------------------------
ModelProvider ModelProvider::with_empty_models() {
    return ModelProvider({std::make_shared<const Model>(Model::empty(ModelType::Acids)),
                          std::make_shared<const Model>(Model::empty(ModelType::QualityScores))});
}

ModelProvider ModelProvider::from_directory(const std::string& dir) {
    DIR* d = opendir(dir.c_str());
    if (!d) throw ModelError("cannot open directory `" + dir + "`");
    std::vector<std::string> names;
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n.size() > 8 && n.compare(n.size() - 8, 8, ".msgpack") == 0) names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    std::vector<std::shared_ptr<const Model>> models;
    for (auto& n : names) models.push_back(std::make_shared<const Model>(Model::read_file(dir + "/" + n)));
    return ModelProvider(std::move(models));
}

std::vector<ModelIdentifier> ModelProvider::identifiers() const {
    std::vector<ModelIdentifier> out;
    for (auto& m : models_) out.push_back(m->identifier());
    return out;
}

size_t ModelProvider::index_of(const ModelIdentifier& id) const {
    for (size_t i = 0; i < models_.size(); i++)
        if (models_[i]->identifier() == id) return i;
    throw ModelError("unknown model " + to_hex(id));
}

bool ModelProvider::has_all_models(const std::vector<ModelIdentifier>& ids) const {
    for (auto& id : ids) {
        bool found = false;
        for (auto& m : models_) found = found || m->identifier() == id;
        if (!found) return false;
    }
    return true;
}

void ModelProvider::filter_by_identifiers(const std::vector<ModelIdentifier>& ids) {
    std::vector<std::shared_ptr<const Model>> kept;
    for (auto& id : ids) kept.push_back(models_[index_of(id)]);
    models_ = std::move(kept);
}

}  // namespace idencomp
