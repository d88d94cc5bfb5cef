// mixq_plugin.cpp -- MixQPlugin / MixQPluginCreator over the C ABI, plus the C handle
// wrappers (mixq_plugin_*) that let ctypes drive the C++ classes.
//
// Behaviour mirrors the reference plugin (TsinghuaMixQPlugin.cpp) where a caller can observe
// it: plugin type/version strings (:180-181), 7 half/linear inputs + 1 half output (:263-320),
// output dims (:244-261), 12-byte serialisation (:808-820, :227-234), creator field parsing
// (:895-933).  Deliberate differences, all documented in DESIGN.md:
//   * enqueue returns non-zero when a launch fails (the reference always returns 0, :402,752);
//   * the workspace is computed in size_t (the reference overflows an int, :373-377) and is
//     M*K + 2M + 256M bytes instead of max(M*K + 2M + 2KN, 16MN);
//   * no cublasHandle_t is created in initialize() (:792-799) -- nothing here needs cuBLAS;
//   * the creator accepts both the advertised field names mm/mn/mk (:873-875) and the names
//     plugin.py actually sends, m/n/k (:906-918).
#include "mixq_plugin.h"

#include <cstring>
#include <iostream>
#include <new>

using namespace nvinfer1;

namespace openai_triton::plugin {

namespace {
char const* const kPluginVersion{"1"};
char const* const kPluginName{"MixQ"};

template <typename T>
void writeArg(char*& buffer, T const& val) {
    std::memcpy(buffer, &val, sizeof(T));
    buffer += sizeof(T);
}
template <typename T>
void readArg(char const*& buffer, T& val) {
    std::memcpy(&val, buffer, sizeof(T));
    buffer += sizeof(T);
}

// rows = product of all but the last dim (TsinghuaMixQPlugin.cpp:390-394)
int64_t rowsOf(Dims const& d) {
    int64_t m = 1;
    for (int i = 0; i < d.nbDims - 1; ++i) m *= d.d[i];
    return m;
}
}  // namespace

PluginFieldCollection MixQPluginCreator::mFC{};
std::vector<PluginField> MixQPluginCreator::mPluginAttributes;

MixQPlugin::MixQPlugin(int m, int n, int k) : mm(m), mn(n), mk(k) {}

MixQPlugin::MixQPlugin(void const* data, size_t length) : mm(0), mn(0), mk(0) {
    if (data && length >= 3 * sizeof(int)) {
        char const* d = static_cast<char const*>(data);
        readArg(d, mm);
        readArg(d, mn);
        readArg(d, mk);
    }
}

IPluginV2DynamicExt* MixQPlugin::clone() const noexcept {
    auto* p = new (std::nothrow) MixQPlugin(*this);
    if (p) p->setPluginNamespace(mNamespace.c_str());
    return p;
}

DimsExprs MixQPlugin::getOutputDimensions(int32_t outputIndex, DimsExprs const* inputs, int32_t nbInputs,
                                          IExprBuilder& exprBuilder) noexcept {
    // [..., K] x W[N, K/2 halves] -> [..., N]
    (void)outputIndex;
    (void)nbInputs;
    DimsExprs ret;
    ret.nbDims = inputs[0].nbDims;
    for (int i = 0; i < ret.nbDims - 1; ++i) ret.d[i] = inputs[0].d[i];
    ret.d[ret.nbDims - 1] = exprBuilder.constant(inputs[1].d[0]->getConstantValue());
    return ret;
}

bool MixQPlugin::supportsFormatCombination(int32_t pos, PluginTensorDesc const* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) noexcept {
    // activation, int8 weight, scales, fp weight, indices, weight-only weight, its scales, output:
    // every tensor travels typed as linear fp16 (int8/int32 payloads are raw bytes, plugin.py:99-111)
    if (pos < 0 || pos >= nbInputs + nbOutputs || pos > kNbInputs) return false;
    return inOut[pos].type == DataType::kHALF && inOut[pos].format == TensorFormat::kLINEAR;
}

void MixQPlugin::configurePlugin(DynamicPluginTensorDesc const* in, int32_t nbInputs,
                                 DynamicPluginTensorDesc const* out, int32_t nbOutputs) noexcept {
    (void)out;
    (void)nbOutputs;
    if (!in || nbInputs < 2) return;
    int64_t const maxM = rowsOf(in[0].max);
    int64_t const maxK = in[0].max.d[in[0].max.nbDims - 1];
    int64_t const maxN = in[1].max.d[0];
    mWorkspaceMaxSize = mixq_workspace_size(maxM, maxN, maxK);
}

size_t MixQPlugin::getWorkspaceSize(PluginTensorDesc const* inputs, int32_t nbInputs, PluginTensorDesc const* outputs,
                                    int32_t nbOutputs) const noexcept {
    (void)outputs;
    (void)nbOutputs;
    size_t need = 0;
    if (inputs && nbInputs >= 2)
        need = mixq_workspace_size(rowsOf(inputs[0].dims), inputs[1].dims.d[0],
                                   inputs[0].dims.d[inputs[0].dims.nbDims - 1]);
    return need > mWorkspaceMaxSize ? need : mWorkspaceMaxSize;
}

int32_t MixQPlugin::enqueue(PluginTensorDesc const* inputDesc, PluginTensorDesc const* outputDesc,
                            void const* const* inputs, void* const* outputs, void* workspace,
                            cudaStream_t stream) noexcept {
    (void)outputDesc;
    if (!inputDesc || !inputs || !outputs) return MIXQ_ERR_BAD_ARG;
    int64_t const M = rowsOf(inputDesc[0].dims);
    int64_t const K = inputDesc[0].dims.d[inputDesc[0].dims.nbDims - 1];
    int64_t const N = inputDesc[1].dims.d[0];
    mixq_tensors t;
    t.A = inputs[0];
    t.W8 = inputs[1];
    t.scale_b = inputs[2];
    t.fp_weight = inputs[3];
    t.ind = inputs[4];
    t.q_weight = inputs[5];
    t.scaling_factors = inputs[6];
    t.Out = outputs[0];
    // TensorRT hands over at least getWorkspaceSize() bytes; that is the only size we can assume.
    size_t const need = mixq_workspace_size(M, N, K);
    size_t const have = mWorkspaceMaxSize > need ? mWorkspaceMaxSize : need;
    return mixq_enqueue(&t, M, N, K, workspace, have, mFlags, stream);
}

DataType MixQPlugin::getOutputDataType(int32_t index, DataType const* inputTypes, int32_t nbInputs) const noexcept {
    (void)index;
    (void)inputTypes;
    (void)nbInputs;
    return DataType::kHALF;
}

char const* MixQPlugin::getPluginType() const noexcept { return kPluginName; }
char const* MixQPlugin::getPluginVersion() const noexcept { return kPluginVersion; }
int32_t MixQPlugin::getNbOutputs() const noexcept { return 1; }
int32_t MixQPlugin::initialize() noexcept { return 0; }
void MixQPlugin::terminate() noexcept {}
size_t MixQPlugin::getSerializationSize() const noexcept { return sizeof(mm) + sizeof(mn) + sizeof(mk); }
void MixQPlugin::serialize(void* buffer) const noexcept {
    char* d = static_cast<char*>(buffer);
    writeArg(d, mm);
    writeArg(d, mn);
    writeArg(d, mk);
}
void MixQPlugin::destroy() noexcept { delete this; }
void MixQPlugin::setPluginNamespace(char const* ns) noexcept { mNamespace = ns ? ns : ""; }
char const* MixQPlugin::getPluginNamespace() const noexcept { return mNamespace.c_str(); }

// --------------------------------------------------------------------------- creator
MixQPluginCreator::MixQPluginCreator() {
    mPluginAttributes.clear();
    mPluginAttributes.emplace_back(PluginField("mm", nullptr, PluginFieldType::kINT32, -1));
    mPluginAttributes.emplace_back(PluginField("mn", nullptr, PluginFieldType::kINT32, -1));
    mPluginAttributes.emplace_back(PluginField("mk", nullptr, PluginFieldType::kINT32, -1));
    mFC.nbFields = static_cast<int32_t>(mPluginAttributes.size());
    mFC.fields = mPluginAttributes.data();
}
char const* MixQPluginCreator::getPluginName() const noexcept { return kPluginName; }
char const* MixQPluginCreator::getPluginVersion() const noexcept { return kPluginVersion; }
PluginFieldCollection const* MixQPluginCreator::getFieldNames() noexcept { return &mFC; }

IPluginV2* MixQPluginCreator::createPlugin(char const* name, PluginFieldCollection const* fc) noexcept {
    (void)name;
    int m = 0, n = 0, k = 0;
    if (fc) {
        for (int i = 0; i < fc->nbFields; ++i) {
            PluginField const& f = fc->fields[i];
            if (!f.name || !f.data || f.type != PluginFieldType::kINT32) continue;
            int const v = *static_cast<int const*>(f.data);
            if (!std::strcmp(f.name, "m") || !std::strcmp(f.name, "mm")) m = v;
            else if (!std::strcmp(f.name, "n") || !std::strcmp(f.name, "mn")) n = v;
            else if (!std::strcmp(f.name, "k") || !std::strcmp(f.name, "mk")) k = v;
        }
    }
    auto* obj = new (std::nothrow) MixQPlugin(m, n, k);
    if (obj) obj->setPluginNamespace(mNamespace.c_str());
    return obj;
}

IPluginV2* MixQPluginCreator::deserializePlugin(char const* name, void const* serialData, size_t serialLength) noexcept {
    (void)name;
    if (!serialData || serialLength < 3 * sizeof(int)) return nullptr;
    auto* obj = new (std::nothrow) MixQPlugin(serialData, serialLength);
    if (obj) obj->setPluginNamespace(mNamespace.c_str());
    return obj;
}
void MixQPluginCreator::setPluginNamespace(char const* ns) noexcept { mNamespace = ns ? ns : ""; }
char const* MixQPluginCreator::getPluginNamespace() const noexcept { return mNamespace.c_str(); }

}  // namespace openai_triton::plugin

// --------------------------------------------------------------------------- C handles
using openai_triton::plugin::MixQPlugin;

struct mixq_plugin_s {
    MixQPlugin* p;
};

namespace {
mixq_plugin_t* wrap(nvinfer1::IPluginV2* base) {
    if (!base) return nullptr;
    auto* h = new (std::nothrow) mixq_plugin_s{static_cast<MixQPlugin*>(base)};
    if (!h) {
        base->destroy();
        return nullptr;
    }
    h->p->initialize();
    return h;
}
nvinfer1::IPluginCreator* find_creator(const char* ns) {
    auto* reg = getPluginRegistry();
    return reg ? reg->getPluginCreator("MixQ", "1", ns ? ns : "") : nullptr;
}
Dims make_dims(const int64_t* d, int nb) {
    Dims r{};
    r.nbDims = nb < 0 ? 0 : (nb > Dims::MAX_DIMS ? Dims::MAX_DIMS : nb);
    for (int i = 0; i < r.nbDims; ++i) r.d[i] = d[i];
    return r;
}
}  // namespace

extern "C" {

mixq_plugin_t* mixq_plugin_create(const char* ns, int m, int n, int k) {
    auto* c = find_creator(ns);
    if (!c) return nullptr;
    PluginField f[3] = {PluginField("m", &m, PluginFieldType::kINT32, 1), PluginField("n", &n, PluginFieldType::kINT32, 1),
                        PluginField("k", &k, PluginFieldType::kINT32, 1)};
    PluginFieldCollection fc;
    fc.nbFields = 3;
    fc.fields = f;
    return wrap(c->createPlugin("tsinghua_mixQ", &fc));  // layer name used by plugin.py:69
}

mixq_plugin_t* mixq_plugin_deserialize(const char* ns, const void* data, size_t len) {
    auto* c = find_creator(ns);
    return c ? wrap(c->deserializePlugin("tsinghua_mixQ", data, len)) : nullptr;
}

mixq_plugin_t* mixq_plugin_clone(const mixq_plugin_t* h) { return h ? wrap(h->p->clone()) : nullptr; }

void mixq_plugin_destroy(mixq_plugin_t* h) {
    if (!h) return;
    h->p->terminate();
    h->p->destroy();
    delete h;
}
const char* mixq_plugin_type(const mixq_plugin_t* h) { return h->p->getPluginType(); }
const char* mixq_plugin_version(const mixq_plugin_t* h) { return h->p->getPluginVersion(); }
const char* mixq_plugin_namespace(const mixq_plugin_t* h) { return h->p->getPluginNamespace(); }
int mixq_plugin_nb_outputs(const mixq_plugin_t* h) { return h->p->getNbOutputs(); }
size_t mixq_plugin_serialization_size(const mixq_plugin_t* h) { return h->p->getSerializationSize(); }
void mixq_plugin_serialize(const mixq_plugin_t* h, void* buffer) { h->p->serialize(buffer); }

int mixq_plugin_supports_format(const mixq_plugin_t* h, int pos, int dtype_code, int format_code) {
    PluginTensorDesc d[MixQPlugin::kNbInputs + 1];
    for (auto& x : d) {
        x = PluginTensorDesc{};
        x.type = DataType::kHALF;
        x.format = TensorFormat::kLINEAR;
    }
    if (pos < 0 || pos > MixQPlugin::kNbInputs) return 0;
    d[pos].type = static_cast<DataType>(dtype_code);
    d[pos].format = static_cast<TensorFormat>(format_code);
    return h->p->supportsFormatCombination(pos, d, MixQPlugin::kNbInputs, 1) ? 1 : 0;
}

size_t mixq_plugin_workspace_size(mixq_plugin_t* h, const int64_t* a_max_dims, int a_nb_dims, int64_t n) {
    DynamicPluginTensorDesc in[2] = {};
    in[0].max = in[0].min = in[0].opt = in[0].desc.dims = make_dims(a_max_dims, a_nb_dims);
    int64_t wd[2] = {n, a_nb_dims > 0 ? a_max_dims[a_nb_dims - 1] / 2 : 0};
    in[1].max = in[1].min = in[1].opt = in[1].desc.dims = make_dims(wd, 2);
    h->p->configurePlugin(in, 2, nullptr, 0);
    PluginTensorDesc d[2] = {in[0].desc, in[1].desc};
    return h->p->getWorkspaceSize(d, 2, nullptr, 0);
}

int mixq_plugin_enqueue(mixq_plugin_t* h, const int64_t* a_dims, int a_nb_dims, int64_t w_dim0,
                        const void* const* inputs, void* const* outputs, void* workspace, void* stream) {
    if (!h || !a_dims || a_nb_dims < 1) return MIXQ_ERR_BAD_ARG;
    PluginTensorDesc in[MixQPlugin::kNbInputs] = {};
    for (auto& x : in) {
        x.type = DataType::kHALF;
        x.format = TensorFormat::kLINEAR;
    }
    in[0].dims = make_dims(a_dims, a_nb_dims);
    int64_t wd[2] = {w_dim0, a_dims[a_nb_dims - 1] / 2};
    in[1].dims = make_dims(wd, 2);
    PluginTensorDesc out = in[0];
    out.dims.d[out.dims.nbDims - 1] = w_dim0;
    return h->p->enqueue(in, &out, inputs, outputs, workspace, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
