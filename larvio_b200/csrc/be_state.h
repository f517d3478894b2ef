// Back-end (LarVio) device state — see be_pipeline.cu.
#pragma once
#include "lvb_internal.h"
struct LvbBackEnd {
  int dummy;
};
