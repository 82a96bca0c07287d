"""bittoss_b200 — B200-native (sm_100a CUDA) implementation of the sboerwinkle/bittoss
per-tick entity-physics hot path behind a C ABI (include/bittoss_b200.h)."""
from .capi import (BtError, World, build, concat_worlds, device_count, lib, synth_boxstack,  # noqa: F401
                   PHASE_ALL, PHASE_CTRLS, PHASE_FINISH, PHASE_GRAVITY, PHASE_PHYSICS, PHASE_PREP)
from .desc import WorldArrays, state_hash  # noqa: F401
