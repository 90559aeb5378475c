// temporary stubs for planning stages not built yet
#include "pop_internal.h"
void pop_plan_free(pop_handle* h) { popk_astar_free(h); popk_corridor_free(h); popk_qp_free(h); }
extern "C" {
int pop_separation_check_batch(pop_handle* h, const int32_t*, const double*, double, int8_t*) { return pop_fail(h, POP_E_UNSUPPORTED, "separation not built"); }
int pop_replan_cycle(pop_handle* h, const pop_cycle_io*) { return pop_fail(h, POP_E_UNSUPPORTED, "cycle not built"); }
int pop_replan_cycle_resident(pop_handle* h, double, double) { return pop_fail(h, POP_E_UNSUPPORTED, "cycle not built"); }
}
