// The object behind mamg_handle: host hierarchy + (optional) device copy.
#pragma once
#include "hierarchy.h"

namespace mamg { struct DeviceState; void device_state_free(DeviceState*); }

struct mamg_handle_s {
  mamg::Hierarchy H;
  mamg::DeviceState* dev = nullptr;  // created by mamg_to_device (csrc/cuda/device.cu)
};
