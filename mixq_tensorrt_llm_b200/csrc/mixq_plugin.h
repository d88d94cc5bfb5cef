// mixq_plugin.h -- TensorRT plugin adapter over the C ABI (include/mixq_b200.h).
//
// Same class names, namespace, plugin identity ("MixQ", "1") and override set as the
// reference (TsinghuaMixQPlugin.h:34-115), so an engine or network built against
// libtrt_llm_custom_plugins.so finds the same creator.  The body of enqueue() is one call to
// mixq_enqueue(); no cuBLAS handle, no device allocation.
#pragma once
#include <string>
#include <vector>

#include "../../include/mixq/trt_shim.h"
#include "../../include/mixq_b200.h"

namespace openai_triton::plugin {

class MixQPlugin : public nvinfer1::IPluginV2DynamicExt {
public:
    MixQPlugin(int m, int n, int k);
    MixQPlugin(void const* data, size_t length);
    ~MixQPlugin() override = default;

    // IPluginV2DynamicExt
    nvinfer1::IPluginV2DynamicExt* clone() const noexcept override;
    nvinfer1::DimsExprs getOutputDimensions(int32_t outputIndex, nvinfer1::DimsExprs const* inputs, int32_t nbInputs,
                                            nvinfer1::IExprBuilder& exprBuilder) noexcept override;
    bool supportsFormatCombination(int32_t pos, nvinfer1::PluginTensorDesc const* inOut, int32_t nbInputs,
                                   int32_t nbOutputs) noexcept override;
    void configurePlugin(nvinfer1::DynamicPluginTensorDesc const* in, int32_t nbInputs,
                         nvinfer1::DynamicPluginTensorDesc const* out, int32_t nbOutputs) noexcept override;
    size_t getWorkspaceSize(nvinfer1::PluginTensorDesc const* inputs, int32_t nbInputs,
                            nvinfer1::PluginTensorDesc const* outputs, int32_t nbOutputs) const noexcept override;
    int32_t enqueue(nvinfer1::PluginTensorDesc const* inputDesc, nvinfer1::PluginTensorDesc const* outputDesc,
                    void const* const* inputs, void* const* outputs, void* workspace,
                    cudaStream_t stream) noexcept override;

    // IPluginV2Ext
    nvinfer1::DataType getOutputDataType(int32_t index, nvinfer1::DataType const* inputTypes,
                                         int32_t nbInputs) const noexcept override;

    // IPluginV2
    char const* getPluginType() const noexcept override;
    char const* getPluginVersion() const noexcept override;
    int32_t getNbOutputs() const noexcept override;
    int32_t initialize() noexcept override;
    void terminate() noexcept override;
    size_t getSerializationSize() const noexcept override;
    void serialize(void* buffer) const noexcept override;
    void destroy() noexcept override;
    void setPluginNamespace(char const* pluginNamespace) noexcept override;
    char const* getPluginNamespace() const noexcept override;

    // Extra knob (not part of the TensorRT surface): MIXQ_FLAG_* forwarded to mixq_enqueue.
    void setFlags(unsigned flags) noexcept { mFlags = flags; }

    static constexpr int kNbInputs = 7;  // plugin.py:142-150

private:
    std::string mNamespace;
    int mm, mn, mk;                // creator fields; serialised, unused at run time (dims come from the descs)
    size_t mWorkspaceMaxSize = 0;  // set by configurePlugin
    unsigned mFlags = 0;
};

class MixQPluginCreator : public nvinfer1::IPluginCreator {
public:
    MixQPluginCreator();
    char const* getPluginName() const noexcept override;
    char const* getPluginVersion() const noexcept override;
    nvinfer1::PluginFieldCollection const* getFieldNames() noexcept override;
    nvinfer1::IPluginV2* createPlugin(char const* name, nvinfer1::PluginFieldCollection const* fc) noexcept override;
    nvinfer1::IPluginV2* deserializePlugin(char const* name, void const* serialData,
                                           size_t serialLength) noexcept override;
    void setPluginNamespace(char const* pluginNamespace) noexcept override;
    char const* getPluginNamespace() const noexcept override;

private:
    static nvinfer1::PluginFieldCollection mFC;
    static std::vector<nvinfer1::PluginField> mPluginAttributes;
    std::string mNamespace;
};

}  // namespace openai_triton::plugin
