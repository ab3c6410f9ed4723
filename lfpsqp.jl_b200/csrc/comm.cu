// comm.cu -- communicator hooks of the column-sharded large-n mode (NCCL all-reduce of the m x m Gram, the m-vector
// t = sum_g J_g v_g and the packed CG scalars).  Filled in by the multi-GPU step; single GPU = no-ops.
#include "large_state.h"
void comm_allreduce(LargeState &, double *, size_t) {}
void comm_allreduce_scalars(LargeState &, unsigned, unsigned) {}
void comm_allreduce_loop_slot(LargeState &, int, int) {}
void comm_allreduce_loop_slots_cg(LargeState &, int) {}
void comm_release(LargeState &) {}
