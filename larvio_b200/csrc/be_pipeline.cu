#include "be_state.h"
int be_alloc(LvbHandle* h) { h->be = new LvbBackEnd(); return LVB_OK; }
void be_free(LvbHandle* h) { delete h->be; h->be = nullptr; }
extern "C" int lvb_process_features(LvbHandle*, const uint8_t*, const double*, const LvbFeature*, const int*, int, LvbImu*, int*, int, uint8_t*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
extern "C" int lvb_step(LvbHandle*, const uint8_t*, int, const double*, LvbImu*, int*, int, uint8_t*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
extern "C" int lvb_set_initial_state(LvbHandle*, int, double, const double*, const double*, const double*, const double*, const double*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
extern "C" int lvb_get_state(LvbHandle*, int, double*, double*, double*, double*, double*, double*, double*, double*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
extern "C" int lvb_get_states(LvbHandle*, double*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
extern "C" int lvb_get_window(LvbHandle*, int, double*, int, int*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
extern "C" int lvb_get_covariance(LvbHandle*, int, double*, int, int*) { return lvb_set_err(LVB_E_UNSUPPORTED, "back end not built yet"); }
