// mixq_registry.cpp -- `extern "C" bool initOpenAiTritonPlugins(void*, char const*)`, the symbol
// the reference's plugin.py resolves with ctypes (plugin.py:34-43; defined in the reference at
// MixQPlugins.cpp:126-132), and -- in shim builds only -- a small in-process
// nvinfer1::IPluginRegistry so the creator can be looked up the way TensorRT would.
//
// Registration semantics kept from MixQPlugins.cpp:42-90: thread safe, idempotent per
// "<namespace>::<name> version <version>", creators owned by the registry object for the
// life of the process, optional ILogger notified.
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "mixq_plugin.h"

namespace {

#ifndef MIXQ_HAVE_TENSORRT
class LocalPluginRegistry final : public nvinfer1::IPluginRegistry {
public:
    bool registerCreator(nvinfer1::IPluginCreator& creator, char const* const ns) noexcept override {
        std::lock_guard<std::mutex> g(mLock);
        std::string const key = keyOf(ns, creator.getPluginName(), creator.getPluginVersion());
        if (mCreators.count(key)) return false;
        mCreators[key] = &creator;
        return true;
    }
    nvinfer1::IPluginCreator* getPluginCreator(char const* const name, char const* const version,
                                               char const* const ns) noexcept override {
        std::lock_guard<std::mutex> g(mLock);
        auto it = mCreators.find(keyOf(ns, name, version));
        return it == mCreators.end() ? nullptr : it->second;
    }
    bool deregisterCreator(nvinfer1::IPluginCreator const& creator) noexcept override {
        std::lock_guard<std::mutex> g(mLock);
        for (auto it = mCreators.begin(); it != mCreators.end(); ++it)
            if (it->second == &creator) {
                mCreators.erase(it);
                return true;
            }
        return false;
    }

private:
    static std::string keyOf(char const* ns, char const* name, char const* version) {
        return std::string(ns ? ns : "") + "::" + (name ? name : "") + " version " + (version ? version : "");
    }
    std::mutex mLock;
    std::map<std::string, nvinfer1::IPluginCreator*> mCreators;
};
#endif

class CreatorOwner {
public:
    static CreatorOwner& instance() {
        static CreatorOwner o;
        return o;
    }
    template <class CreatorT>
    bool add(void* logger, char const* libNamespace) {
        std::lock_guard<std::mutex> g(mLock);
        auto creator = std::make_unique<CreatorT>();
        creator->setPluginNamespace(libNamespace);
        std::string const id = std::string(creator->getPluginNamespace()) + "::" + creator->getPluginName() +
                               " version " + creator->getPluginVersion();
        auto* trtLogger = static_cast<nvinfer1::ILogger*>(logger);
        bool ok = true;
        std::string msg;
        if (mKnown.count(id)) {
            msg = "Plugin creator already registered - " + id;
        } else if (getPluginRegistry()->registerCreator(*creator, libNamespace)) {
            mKnown[id] = true;
            mOwned.push_back(std::move(creator));
            msg = "Registered plugin creator - " + id;
        } else {
            ok = false;
            msg = "Could not register plugin creator -  " + id;
        }
        if (trtLogger)
            trtLogger->log(ok ? nvinfer1::ILogger::Severity::kVERBOSE : nvinfer1::ILogger::Severity::kERROR, msg.c_str());
        return ok;
    }

private:
    std::mutex mLock;
    std::vector<std::unique_ptr<nvinfer1::IPluginCreator>> mOwned;
    std::map<std::string, bool> mKnown;
};

}  // namespace

#ifndef MIXQ_HAVE_TENSORRT
extern "C" nvinfer1::IPluginRegistry* getPluginRegistry() noexcept {
    static LocalPluginRegistry registry;
    return &registry;
}
#endif

extern "C" bool initOpenAiTritonPlugins(void* logger, char const* libNamespace) {
    // the reference returns true unconditionally (MixQPlugins.cpp:128-131); so do we -- a
    // failed registration is reported through the logger, as there.
    CreatorOwner::instance().add<openai_triton::plugin::MixQPluginCreator>(logger, libNamespace ? libNamespace : "");
    return true;
}
