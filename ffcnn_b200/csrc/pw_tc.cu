/* pw_tc.cu -- placeholder until the tcgen05 kernel lands: no shape is eligible yet. */
#include "pw_tc.h"
PwTcPlan   *pw_tc_plan_create(int, int, int, int) { return nullptr; }
void        pw_tc_plan_destroy(PwTcPlan *) {}
int         pw_tc_prepare(PwTcPlan *, const float *, int, cudaStream_t) { return -1; }
int         pw_tc_run(PwTcPlan *, const float *, int, float *, int, int, long, cudaStream_t) { return -1; }
const char *pw_tc_mode_name(const PwTcPlan *) { return "none"; }
