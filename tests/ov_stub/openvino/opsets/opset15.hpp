// forwards to the stand-in OpenVINO API (test infrastructure, see stub_core.hpp)
#pragma once
#include "openvino/stub_core.hpp"
