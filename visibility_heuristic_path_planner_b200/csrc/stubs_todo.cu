// Temporary: entry points of include/vhp.h that are not implemented yet.
#include "vhp_internal.h"
extern "C" {
vhp_status vhp_planner_batch(vhp_context *, const uint8_t *, int, int, int, const int32_t *,
                             const int32_t *, int64_t, double, int32_t, int32_t, vhp_dtype,
                             const vhp_planner_out *) { return VHP_ERR_UNSUPPORTED; }
vhp_status vhp_planner_batch_dev(vhp_context *, const uint8_t *, int, int, int, const int32_t *,
                                 const int32_t *, int64_t, double, int32_t, int32_t, vhp_dtype,
                                 const vhp_planner_out *) { return VHP_ERR_UNSUPPORTED; }
}
