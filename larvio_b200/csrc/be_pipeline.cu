#include "be_state.h"
int be_alloc(LvbHandle* h) { h->be = new LvbBackEnd(); return LVB_OK; }
void be_free(LvbHandle* h) { delete h->be; h->be = nullptr; }
