// model.hpp -- host-side mirror of idencomp's model types (the drop-in surface above the C-ABI).
//
//   ContextSpecType   context_spec.rs:532-599 + idencomp-macros/src/lib.rs:166-196 (names), :305-319 (spec_num)
//   Model             model.rs:176-259 (contexts sorted by their spec lists, spec -> context map), :458-482 (identifier)
//   ModelProvider     idn/model_provider.rs:51-381
//   quantise()        Context::as_integer_cum_freqs + fix_zero_freqs, context.rs:346-394 (bit-exact f32)
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/idn_gpu.h"

namespace idencomp {

constexpr uint32_t kScaleBits = 14;  // idn/model_provider.rs:407

enum class ModelType : uint8_t { Acids = 0, QualityScores = 1 };  // model.rs:56-61 (repr u8, hashed into the identifier)

inline uint32_t symbols_of(ModelType t) { return t == ModelType::Acids ? 5u : 94u; }

struct ContextSpecType {
    enum Kind { Generic = 0, Light = 1 } kind = Generic;
    int acid_order = 0, q_score_order = 0, position_bits = 0, q_score_max = 0;
    std::string name = "dummy";

    // parses "dummy" | "generic_ao{A}_qo{Q}_pb{P}" | "light_ao{A}_qo{Q}_pb{P}_qm{M}"
    static ContextSpecType parse(const std::string& name);
    uint32_t bits() const;       // acid_bits + q_score_bits + position_bits
    uint64_t spec_num() const;   // 1 << bits()
};

using ModelIdentifier = std::array<uint8_t, 32>;
std::string to_hex(const ModelIdentifier& id);

struct ModelContext {
    std::vector<uint32_t> specs;      // sorted context specs merged into this context (context_binning output)
    float context_prob = 0.f;
    std::vector<float> symbol_prob;   // 5 or 94 entries
};

struct ModelError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class Model {
public:
    // Model::with_model_and_spec_type: sorts contexts by their spec lists, computes the identifier
    Model(ModelType type, ContextSpecType spec_type, std::vector<ModelContext> contexts);
    static Model empty(ModelType type);                    // model.rs:251-259
    static Model read_msgpack(const uint8_t* data, size_t n);  // model_serializer.rs:90-114 (asserts the identifier)
    static Model read_file(const std::string& path);

    ModelType model_type() const { return type_; }
    const ContextSpecType& context_spec_type() const { return spec_; }
    const ModelIdentifier& identifier() const { return id_; }
    size_t len() const { return contexts_.size(); }
    const std::vector<ModelContext>& contexts() const { return contexts_; }

    // integer tables the C-ABI takes: row 0 = Context::dummy (sequence_compressor.rs:26-29), rows 1.. = contexts
    std::vector<uint16_t> cum_table() const;  // [(len+1)][nsym+1]
    void spec_table(std::vector<uint32_t>* keys, std::vector<uint32_t>* ctx) const;

private:
    ModelType type_;
    ContextSpecType spec_;
    std::vector<ModelContext> contexts_;
    ModelIdentifier id_{};
    void make_identifier();
};

// Context::as_integer_cum_freqs (context.rs:346-371): cum[nsym] exclusive prefix sums, all freqs >= 1, total 2^scale_bits
std::vector<uint32_t> quantise(const float* probs, size_t nsym, uint32_t scale_bits);

// RansEncModel/RansDecModel::from_model on the device: integer tables -> idn_gpu_model_upload
int32_t upload_model(idn_gpu_ctx* ctx, const Model& m, idn_model_t* handle);

class ModelProvider {
public:
    ModelProvider() = default;
    explicit ModelProvider(std::vector<std::shared_ptr<const Model>> models) : models_(std::move(models)) {}
    static ModelProvider with_empty_models();  // one empty model per type (model_provider.rs:67-73)
    // the reference iterates fs::read_dir order (filesystem dependent); here: file names sorted
    static ModelProvider from_directory(const std::string& dir);

    size_t len() const { return models_.size(); }
    const Model& operator[](size_t i) const { return *models_[i]; }
    std::shared_ptr<const Model> ptr(size_t i) const { return models_[i]; }
    std::vector<ModelIdentifier> identifiers() const;
    // position of the model in this provider; throws ModelError if absent (the reference panics)
    size_t index_of(const ModelIdentifier& id) const;
    bool has_all_models(const std::vector<ModelIdentifier>& ids) const;
    // keep exactly `ids`, in that order (model_provider.rs:290-329); throws ModelError on an unknown identifier
    void filter_by_identifiers(const std::vector<ModelIdentifier>& ids);

private:
    std::vector<std::shared_ptr<const Model>> models_;
};

}  // namespace idencomp
