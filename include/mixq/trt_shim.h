/*
 * trt_shim.h -- the slice of <NvInferRuntime.h> (TensorRT 10.x, the version TensorRT-LLM
 * 0.12.0.dev builds against) that the MixQ plugin touches, declared with the same names,
 * enumerator values and member layouts so that mixq_plugin.{h,cpp} compiles unchanged
 * against either this file or the real header.
 *
 * TensorRT is not installed in the build image, so this mirror is what the always-built
 * library and the tests use; when <NvInferRuntime.h> is on the include path
 * (-DMIXQ_USE_TENSORRT or automatic via __has_include) the real header is used instead and
 * the plugin registers with TensorRT's own registry.
 *
 * Reference anchors: TsinghuaMixQPlugin.h:34-115 (which virtuals are overridden),
 * TsinghuaMixQPlugin.cpp:263-320 (PluginTensorDesc use), :325-349 (DynamicPluginTensorDesc),
 * :895-933 (PluginFieldCollection), MixQPlugins.cpp:55-76 (registry use).
 */
#pragma once

#if !defined(MIXQ_FORCE_TRT_SHIM) && defined(__has_include)
#if __has_include(<NvInferRuntime.h>)
#define MIXQ_HAVE_TENSORRT 1
#endif
#endif

#ifdef MIXQ_HAVE_TENSORRT
#include <NvInferRuntime.h>
#else

#include <cuda_runtime_api.h>

#include <cstddef>
#include <cstdint>

namespace nvinfer1 {

using AsciiChar = char;

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4, kUINT8 = 5, kFP8 = 6, kBF16 = 7, kINT64 = 8, kINT4 = 9 };
enum class TensorFormat : int32_t { kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3, kCHW16 = 4, kCHW32 = 5 };
using PluginFormat = TensorFormat;

class Dims64 {
public:
    static constexpr int32_t MAX_DIMS{8};
    int32_t nbDims;
    int64_t d[MAX_DIMS];
};
using Dims = Dims64;

struct PluginTensorDesc {
    Dims dims;
    DataType type;
    TensorFormat format;
    float scale;
};

struct DynamicPluginTensorDesc {
    PluginTensorDesc desc;
    Dims min;
    Dims max;
    Dims opt;
};

enum class DimensionOperation : int32_t { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8 };

class IDimensionExpr {
public:
    virtual bool isConstant() const noexcept = 0;
    virtual int64_t getConstantValue() const noexcept = 0;
    virtual bool isSizeTensor() const noexcept { return false; }
protected:
    virtual ~IDimensionExpr() noexcept = default;
};

class DimsExprs {
public:
    int32_t nbDims;
    IDimensionExpr const* d[Dims::MAX_DIMS];
};

class IExprBuilder {
public:
    virtual IDimensionExpr const* constant(int64_t value) noexcept = 0;
    virtual IDimensionExpr const* operation(DimensionOperation op, IDimensionExpr const& first,
                                            IDimensionExpr const& second) noexcept = 0;
protected:
    virtual ~IExprBuilder() noexcept = default;
};

enum class PluginFieldType : int32_t { kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8 };

class PluginField {
public:
    AsciiChar const* name;
    void const* data;
    PluginFieldType type;
    int32_t length;
    PluginField(AsciiChar const* const name_ = nullptr, void const* const data_ = nullptr,
                PluginFieldType const type_ = PluginFieldType::kUNKNOWN, int32_t const length_ = 0) noexcept
        : name(name_), data(data_), type(type_), length(length_) {}
};

struct PluginFieldCollection {
    int32_t nbFields{};
    PluginField const* fields{};
};

class ILogger {
public:
    enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    virtual void log(Severity severity, AsciiChar const* msg) noexcept = 0;
    virtual ~ILogger() = default;
};

class IPluginV2 {
public:
    virtual AsciiChar const* getPluginType() const noexcept = 0;
    virtual AsciiChar const* getPluginVersion() const noexcept = 0;
    virtual int32_t getNbOutputs() const noexcept = 0;
    virtual int32_t initialize() noexcept = 0;
    virtual void terminate() noexcept = 0;
    virtual size_t getSerializationSize() const noexcept = 0;
    virtual void serialize(void* buffer) const noexcept = 0;
    virtual void destroy() noexcept = 0;
    virtual void setPluginNamespace(AsciiChar const* pluginNamespace) noexcept = 0;
    virtual AsciiChar const* getPluginNamespace() const noexcept = 0;
    virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2 {
public:
    virtual DataType getOutputDataType(int32_t index, DataType const* inputTypes, int32_t nbInputs) const noexcept = 0;
};

class IPluginV2DynamicExt : public IPluginV2Ext {
public:
    virtual IPluginV2DynamicExt* clone() const noexcept = 0;
    virtual DimsExprs getOutputDimensions(int32_t outputIndex, DimsExprs const* inputs, int32_t nbInputs,
                                          IExprBuilder& exprBuilder) noexcept = 0;
    virtual bool supportsFormatCombination(int32_t pos, PluginTensorDesc const* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) noexcept = 0;
    virtual void configurePlugin(DynamicPluginTensorDesc const* in, int32_t nbInputs,
                                 DynamicPluginTensorDesc const* out, int32_t nbOutputs) noexcept = 0;
    virtual size_t getWorkspaceSize(PluginTensorDesc const* inputs, int32_t nbInputs, PluginTensorDesc const* outputs,
                                    int32_t nbOutputs) const noexcept = 0;
    virtual int32_t enqueue(PluginTensorDesc const* inputDesc, PluginTensorDesc const* outputDesc,
                            void const* const* inputs, void* const* outputs, void* workspace,
                            cudaStream_t stream) noexcept = 0;
};

class IPluginCreator {
public:
    virtual AsciiChar const* getPluginName() const noexcept = 0;
    virtual AsciiChar const* getPluginVersion() const noexcept = 0;
    virtual PluginFieldCollection const* getFieldNames() noexcept = 0;
    virtual IPluginV2* createPlugin(AsciiChar const* name, PluginFieldCollection const* fc) noexcept = 0;
    virtual IPluginV2* deserializePlugin(AsciiChar const* name, void const* serialData, size_t serialLength) noexcept = 0;
    virtual void setPluginNamespace(AsciiChar const* pluginNamespace) noexcept = 0;
    virtual AsciiChar const* getPluginNamespace() const noexcept = 0;
    virtual ~IPluginCreator() = default;
};

class IPluginRegistry {
public:
    virtual bool registerCreator(IPluginCreator& creator, AsciiChar const* const pluginNamespace) noexcept = 0;
    virtual IPluginCreator* getPluginCreator(AsciiChar const* const pluginName, AsciiChar const* const pluginVersion,
                                             AsciiChar const* const pluginNamespace = "") noexcept = 0;
    virtual bool deregisterCreator(IPluginCreator const& creator) noexcept = 0;
    virtual ~IPluginRegistry() noexcept = default;
};

}  // namespace nvinfer1

// Provided by mixq_registry.cpp in shim builds; by libnvinfer otherwise.
extern "C" nvinfer1::IPluginRegistry* getPluginRegistry() noexcept;

#endif  // MIXQ_HAVE_TENSORRT
