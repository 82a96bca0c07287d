/* bt_small.h — launcher of the one-block-per-world step (bt_small.cu), called by bt_world.cu */
#ifndef BT_SMALL_H
#define BT_SMALL_H
#include "bt_types.h"

#define SW_PHASES 128               /* root-order phases (rootKey, bt_fumble) one launch may hand out: the opening flush1, one
                                     * per doIteration that pushes, the closing flush1 */
#define SWF_HOST_DELTAS 0x80000000u /* in `phases`: d_center / d_vel may be non-zero on entry (host writes, a gravity-only step) */

/* one block per world of W; one object file per (block size, resident warps per SM) variant of bt_small.cu */
cudaError_t bt_launch_tick_small_64_16(const DevWorld &W, unsigned phases, bool pdl, cudaStream_t stream);
cudaError_t bt_launch_tick_small_64_32(const DevWorld &W, unsigned phases, bool pdl, cudaStream_t stream);
cudaError_t bt_launch_tick_small_128_16(const DevWorld &W, unsigned phases, bool pdl, cudaStream_t stream);
cudaError_t bt_launch_tick_small_128_32(const DevWorld &W, unsigned phases, bool pdl, cudaStream_t stream);
/* block = 64 threads when no world holds more than 64 slots, else BT_SMALL_WORLD; the 32-warp variant (64 registers,
 * spills) once the worlds fill every SM at that occupancy, the 16-warp one (128 registers) below */
static inline cudaError_t bt_launch_tick_small(const DevWorld &W, unsigned phases, int biggest_world, int warps, bool pdl, cudaStream_t stream) {
	const int bs = biggest_world <= 64 ? 64 : BT_SMALL_WORLD;
	const bool many = warps ? warps >= 32 : (long long)W.n_worlds * (bs / 32) >= 148LL * 32; /* warps: option small_warps (0 = by the number of worlds) */
	if (bs == 64) return many ? bt_launch_tick_small_64_32(W, phases, pdl, stream) : bt_launch_tick_small_64_16(W, phases, pdl, stream);
	return many ? bt_launch_tick_small_128_32(W, phases, pdl, stream) : bt_launch_tick_small_128_16(W, phases, pdl, stream);
}

#endif
