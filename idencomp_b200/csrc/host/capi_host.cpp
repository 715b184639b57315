// capi_host.cpp -- flat C view (include/idn_host.h) of the C++ host mirror.
#include "../../../include/idn_host.h"

#include <cstring>
#include <string>

#include "model.hpp"

using namespace idencomp;

struct idn_host_model {
    std::shared_ptr<const Model> m;
};

namespace {
thread_local std::string g_err;

int32_t set_err(int32_t code, const std::string& what) {
    g_err = what;
    return code;
}

template <class F>
int32_t guarded(F&& f) {
    try {
        return f();
    } catch (const ModelError& e) {
        return set_err(IDN_E_SERIALIZE, e.what());
    } catch (const std::exception& e) {
        return set_err(IDN_E_INVALID_STATE, e.what());
    }
}
}  // namespace

extern "C" const char* idn_host_last_error(void) { return g_err.c_str(); }

extern "C" int32_t idn_host_model_load(const char* path, idn_host_model** out) {
    if (!path || !out) return set_err(IDN_E_INVALID_ARG, "NULL argument");
    *out = nullptr;
    return guarded([&] {
        *out = new idn_host_model{std::make_shared<const Model>(Model::read_file(path))};
        return (int32_t)IDN_OK;
    });
}

extern "C" int32_t idn_host_model_from_bytes(const uint8_t* data, size_t n, idn_host_model** out) {
    if (!data || !out) return set_err(IDN_E_INVALID_ARG, "NULL argument");
    *out = nullptr;
    return guarded([&] {
        *out = new idn_host_model{std::make_shared<const Model>(Model::read_msgpack(data, n))};
        return (int32_t)IDN_OK;
    });
}

extern "C" int32_t idn_host_model_new(int32_t model_type, const char* spec_name, uint32_t n_ctx, const float* probs,
                                      const uint32_t* spec_keys, const uint32_t* spec_ctx, uint64_t n_specs,
                                      idn_host_model** out) {
    if (!spec_name || !out || (n_ctx && !probs) || (n_specs && (!spec_keys || !spec_ctx)))
        return set_err(IDN_E_INVALID_ARG, "NULL argument");
    if (model_type != IDN_MODEL_ACID && model_type != IDN_MODEL_QSCORE) return set_err(IDN_E_INVALID_ARG, "bad model type");
    *out = nullptr;
    return guarded([&] {
        ModelType t = (ModelType)model_type;
        size_t nsym = symbols_of(t);
        std::vector<ModelContext> ctxs(n_ctx);
        for (uint32_t i = 0; i < n_ctx; i++) ctxs[i].symbol_prob.assign(probs + i * nsym, probs + (i + 1) * nsym);
        for (uint64_t i = 0; i < n_specs; i++) {
            if (spec_ctx[i] >= n_ctx) throw ModelError("context index out of range");
            ctxs[spec_ctx[i]].specs.push_back(spec_keys[i]);
        }
        *out = new idn_host_model{std::make_shared<const Model>(Model(t, ContextSpecType::parse(spec_name), std::move(ctxs)))};
        return (int32_t)IDN_OK;
    });
}

extern "C" int32_t idn_host_model_empty(int32_t model_type, idn_host_model** out) {
    if (!out) return set_err(IDN_E_INVALID_ARG, "NULL argument");
    return guarded([&] {
        *out = new idn_host_model{std::make_shared<const Model>(Model::empty((ModelType)model_type))};
        return (int32_t)IDN_OK;
    });
}

extern "C" void idn_host_model_free(idn_host_model* m) { delete m; }
extern "C" int32_t idn_host_model_type(const idn_host_model* m) { return (int32_t)m->m->model_type(); }
extern "C" uint32_t idn_host_model_len(const idn_host_model* m) { return (uint32_t)m->m->len(); }
extern "C" const char* idn_host_model_spec_name(const idn_host_model* m) { return m->m->context_spec_type().name.c_str(); }
extern "C" void idn_host_model_identifier(const idn_host_model* m, uint8_t out[32]) {
    std::memcpy(out, m->m->identifier().data(), 32);
}

extern "C" uint64_t idn_host_model_cum_table(const idn_host_model* m, uint16_t* out, uint64_t cap) {
    uint64_t need = (uint64_t)(m->m->len() + 1) * (symbols_of(m->m->model_type()) + 1);
    if (!out || cap < need) return need;
    std::vector<uint16_t> t = m->m->cum_table();
    std::memcpy(out, t.data(), t.size() * 2);
    return need;
}

extern "C" int32_t idn_host_model_upload(idn_gpu_ctx* ctx, const idn_host_model* m, idn_model_t* handle) {
    if (!ctx || !m || !handle) return set_err(IDN_E_INVALID_ARG, "NULL argument");
    return guarded([&] {
        int32_t rc = upload_model(ctx, *m->m, handle);
        if (rc) g_err = idn_gpu_last_error(ctx);
        return rc;
    });
}

extern "C" int32_t idn_host_quantise(const float* probs, uint32_t nsym, uint32_t scale_bits, uint32_t* cum_out) {
    if (!probs || !cum_out) return set_err(IDN_E_INVALID_ARG, "NULL argument");
    return guarded([&] {
        std::vector<uint32_t> c = quantise(probs, nsym, scale_bits);
        std::memcpy(cum_out, c.data(), c.size() * 4);
        return (int32_t)IDN_OK;
    });
}
