/*
 * bt_synth.cpp — host-side builder of the synthetic box-stack world of
 * BASELINE.json configs[2] (SURVEY.md §8d "Config 3"): ground tiles + columns of
 * `height` cubes with seeded jitter, a few top cubes kicked sideways.  Fills plain
 * bt_world_desc arrays; the same recipe, built through the reference's own initEnt, lives
 * in oracle/ref_harness.cpp (btref_make_boxstack) and tests check both agree.
 * RNG: splitmix32 as in random.cpp:18-24.
 */
#include <stdint.h>
#include <string.h>

#include "../../include/bittoss_b200.h"

static uint32_t splitmix32(uint32_t *state) {
	uint32_t z = (*state += 0x9e3779b9);
	z ^= z >> 16; z *= 0x21f0aaad;
	z ^= z >> 15; z *= 0x735a2d97;
	z ^= z >> 15;
	return z;
}

static void layout(int32_t n_cubes, int32_t H, int32_t *n_cols, int32_t *G, int32_t *span, int32_t *T) {
	*n_cols = (n_cubes + H - 1) / H;
	int32_t g = 1;
	while ((int64_t)g * g < *n_cols) g++;
	*G = g;
	*span = g * 1000;
	*T = (*span + 6400 + 6399) / 6400;
}

extern "C" int bt_synth_boxstack_size(int32_t n_cubes, int32_t height, int32_t *n_ents) {
	if (n_cubes < 0 || height < 1 || !n_ents) return BT_ERR_ARG;
	int32_t n_cols, G, span, T;
	layout(n_cubes, height, &n_cols, &G, &span, &T);
	*n_ents = T * T + n_cubes;
	return BT_OK;
}

extern "C" int bt_synth_boxstack_fill(bt_world_desc *d, int32_t n_cubes, int32_t H, uint32_t seed, int32_t push_every,
                                      int32_t push_max) {
	if (!d || n_cubes < 0 || H < 1) return BT_ERR_ARG;
	int32_t n_cols, G, span, T;
	layout(n_cubes, H, &n_cols, &G, &span, &T);
	const int32_t n = T * T + n_cubes;
	if (d->n_ents < n) return BT_ERR_CAPACITY;
	d->n_ents = n; d->n_sliders = 0; d->n_wires = 0; d->n_worlds = 1;
	d->w_flip_flop[0] = 0; d->w_gamerules[0] = 1 << 3 /* EFFECT_GRAV effects.h:11 */; d->w_rand[0] = 12345;
	int32_t i = 0;
	auto put = [&](const int32_t c[3], const int32_t v[3], const int32_t r[3], uint32_t type, uint32_t collide,
	               uint8_t whoMoves) {
		if (d->world_id) d->world_id[i] = 0;
		d->uid[i] = i;
		for (int k = 0; k < 3; k++) {
			d->center[3 * i + k] = c[k]; d->vel[3 * i + k] = v[k]; d->old[3 * i + k] = c[k] - v[k]; /* entFuncs.c:83 */
			d->d_center[3 * i + k] = 0; d->d_vel[3 * i + k] = 0; d->radius[3 * i + k] = r[k];
		}
		d->type_mask[i] = d->new_type_mask[i] = type;
		d->collide_mask[i] = d->new_collide_mask[i] = collide;
		d->friction[i] = 16; /* DEFAULT_FRICTION entFuncs.h:1 */
		d->holder[i] = -1; d->hold_flags[i] = 0; d->holdee_head[i] = -1; d->holdee_next[i] = -1;
		d->root_order[i] = n - 1 - i; /* rootEnts is newest-first (ent.cpp:60-66) */
		d->new_holder[i] = -1; d->new_hold_flags[i] = 0;
		memset(d->handlers + (size_t)BT_H_NUM * i, 0, BT_H_NUM);
		d->handlers[(size_t)BT_H_NUM * i + BT_H_WHOMOVES] = whoMoves;
		memset(d->btn + 8 * (size_t)i, 0, 8);
		memset(d->misc + (size_t)BT_M_NUM * i, 0, BT_M_NUM);
		d->slider_off[i] = 0; d->wire_off[i] = 0;
		i++;
	};
	const int32_t zero[3] = {0, 0, 0}, tile_r[3] = {3200, 3200, 500}, cube_r[3] = {400, 400, 400};
	for (int tx = 0; tx < T; tx++)
		for (int ty = 0; ty < T; ty++) {
			int32_t c[3] = {tx * 6400 - (T * 6400) / 2 + 3200, ty * 6400 - (T * 6400) / 2 + 3200, 600};
			put(c, zero, tile_r, 2 | 1 | 16 /* TERRAIN|HEAVY|WEIGHTLESS */, 0, 1 /* WHOMOVES_ME */);
		}
	uint32_t st = seed;
	int32_t made = 0;
	for (int32_t col = 0; col < n_cols; col++) {
		int32_t gx = col / G, gy = col % G;
		int32_t cx = gx * 1000 - span / 2 + 500, cy = gy * 1000 - span / 2 + 500;
		for (int32_t k = 0; k < H && made < n_cubes; k++, made++) {
			int32_t jx = (int32_t)(splitmix32(&st) % 101u) - 50;
			int32_t jy = (int32_t)(splitmix32(&st) % 101u) - 50;
			int32_t jz = (int32_t)(splitmix32(&st) % 64u);
			int32_t c[3] = {cx + jx, cy + jy, 100 - 400 - 900 * k - jz};
			int32_t v[3] = {0, 0, 0};
			if (k == H - 1 && push_every > 0 && col % push_every == 0) {
				v[0] = (int32_t)(splitmix32(&st) % (uint32_t)(2 * push_max + 1)) - push_max;
				v[1] = (int32_t)(splitmix32(&st) % (uint32_t)(2 * push_max + 1)) - push_max;
			}
			put(c, v, cube_r, 4 | 1 /* OBSTACLE|HEAVY */, 2 | 4 /* TERRAIN|OBSTACLE */, 6 /* WHOMOVES_WOOD */);
		}
	}
	d->slider_off[n] = 0; d->wire_off[n] = 0;
	return BT_OK;
}
