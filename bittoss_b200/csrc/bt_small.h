/* bt_small.h — launcher of the one-block-per-world step (bt_small.cu), called by bt_world.cu */
#ifndef BT_SMALL_H
#define BT_SMALL_H
#include "bt_types.h"

#define SW_PHASES 128               /* root-order phases (rootKey, bt_fumble) one launch may hand out: the opening flush1, one
                                     * per doIteration that pushes, the closing flush1 */
#define SWF_HOST_DELTAS 0x80000000u /* in `phases`: d_center / d_vel may be non-zero on entry (host writes, a gravity-only step) */

/* one block per world of W; block = 64 threads when no world holds more than 64 slots, else BT_SMALL_WORLD */
cudaError_t bt_launch_tick_small(const DevWorld &W, unsigned phases, int biggest_world, bool pdl, cudaStream_t stream);

#endif
