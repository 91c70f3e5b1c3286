// shim_driver.cpp — drives this repo's ov::Op shim (openvino_tokenizers_b200/csrc/ov_shim/ov_extension_b200.cpp) through the
// stand-in OpenVINO runtime: the ops are taken from the extension entry point the real runtime calls (create_extensions, generated
// by OPENVINO_CREATE_EXTENSIONS), layers are "loaded" with OpExtension::create, evaluate() runs on host tensors.
// TEST INFRASTRUCTURE ONLY; built by tests/shimlib.py together with the shim into ov_shim/libb200tok_ov_stub.so.
#include <stub_driver.hpp>

extern "C" void create_extensions(std::vector<ov::Extension::Ptr>&);

namespace {
const ovs::Registry& registry() {
    static const ovs::Registry reg = [] {
        ovs::Registry r;
        std::vector<ov::Extension::Ptr> ext;
        create_extensions(ext);
        for (auto& e : ext)
            if (auto op = std::dynamic_pointer_cast<ov::BaseOpExtension>(e)) r[op->get_type_info().name] = op;
        return r;
    }();
    return reg;
}
}  // namespace
OVS_DEFINE_C_API(ovshim, registry())

// a chain of layers loaded one after the other, each fed from the previous one's outputs — what reading an IR does; used to
// check the load-time fusion of RegexSplit -> BPETokenizer / WordpieceTokenizer
extern "C" __attribute__((visibility("default"))) int ovshim_registered(const char* op) { return registry().count(op) ? 1 : 0; }
