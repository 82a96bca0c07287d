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
 * spills, state in global memory) once the worlds fill every SM twice over at that occupancy, the 16-warp one (128
 * registers, state staged in shared memory) below: measured on arena.sav worlds, 4096 of them 0.081 (16) / 0.084 ms
 * (32) per tick, 8192 of them 0.191 / 0.173 */
static inline cudaError_t bt_launch_tick_small(const DevWorld &W, unsigned phases, int biggest_world, int warps, bool pdl, cudaStream_t stream) {
	const int bs = biggest_world <= 64 ? 64 : BT_SMALL_WORLD;
	const bool many = warps ? warps >= 32 : (long long)W.n_worlds * (bs / 32) >= 148LL * 32 * 2; /* warps: option small_warps (0 = by the number of worlds) */
	if (bs == 64) return many ? bt_launch_tick_small_64_32(W, phases, pdl, stream) : bt_launch_tick_small_64_16(W, phases, pdl, stream);
	return many ? bt_launch_tick_small_128_32(W, phases, pdl, stream) : bt_launch_tick_small_128_16(W, phases, pdl, stream);
}

#endif
