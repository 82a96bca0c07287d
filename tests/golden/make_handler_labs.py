"""Hand-built "handler lab" worlds: each exercises collision-path handlers that no kicked shipped
level triggers without players (PUSH_TELEPORT / SEAT / PUDDLEJUMP / BOTTLE, PICKUP_SEAT,
FUMBLE_SEAT(_OLD), crush handlers).  Built through the UNMODIFIED reference (initEnt, its handler
catalogs and deferred updaters) via oracle/_ref, stepped physics-only (prepPhysics, doGravity,
doPhysics, finishStep: main.cpp:1076-1096) and saved as fixtures:
initial world + per-tick state hashes + final full state (+ the order crush handlers ran in).

    python tests/golden/make_handler_labs.py          # needs /root/reference (oracle/_ref built)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

T_HEAVY, T_TERRAIN, T_OBSTACLE, T_DEBRIS, T_WEIGHTLESS, T_FLAG = 1, 2, 4, 8, 16, 32
T_DECOR = 64
T_INPUTS, T_EQUIP, T_EQUIP_SM = 1 << 11, 1 << 12, 1 << 13
HOLD_DROP, HOLD_MOVE, HOLD_FREEZE, HOLD_SINGLE = 1, 2, 4, 8
WM_ME, WM_WOOD = 1, 6
PUSH_SEAT, PUSH_TELEPORT, PUSH_PUDDLEJUMP, PUSH_BOTTLE = 3, 6, 8, 9
PICKUP_SEAT, FUMBLE_SEAT_OLD, FUMBLE_SEAT = 1, 2, 4
TICK_BOTTLE = 37
EFFECT_GRAV = 8

GROUND = T_TERRAIN | T_HEAVY | T_WEIGHTLESS
CUBE = T_OBSTACLE | T_HEAVY
CUBE_COLLIDE = T_TERRAIN | T_OBSTACLE


def pad(R, x, y, r=(1500, 1500, 300), z=400):
    """immovable slab whose top face is at z-300 (+z is down)"""
    return R.add_ent((x, y, z), (0, 0, 0), r, GROUND, 0, WM_ME)


def cube(R, x, y, z, v=(0, 0, 0), r=400, extra_type=0):
    return R.add_ent((x, y, z), v, (r, r, r), CUBE | extra_type, CUBE_COLLIDE, WM_WOOD)


def lab_teleport(R):
    """two pads wired to destination markers; one destination is lit (button 2), cubes rain on both"""
    ev = []
    a = pad(R, 0, 0)
    b = pad(R, 6000, 0)
    c = pad(R, 12000, 0)          # two lit destinations: undefined -> no teleport (teleport.c:19-23)
    da = R.add_ent((40000, 5000, -3000), (7, -3, 0), (10, 10, 10), T_WEIGHTLESS, 0, WM_ME)
    db = R.add_ent((-40000, 0, -1000), (0, 0, 0), (10, 10, 10), T_WEIGHTLESS, 0, WM_ME)
    for p, dests in ((a, (da,)), (b, (db,)), (c, (da, db))):
        R.lab_set_handlers(p, push=PUSH_TELEPORT)
        for d in dests:
            R.lab_add_wire(p, d)
    floor = pad(R, 40000, 5000, r=(4000, 4000, 300), z=2000)   # something to land on after the jump
    for i, (x, y) in enumerate(((-300, 200), (500, -400), (6100, 100), (12100, -50))):
        cube(R, x, y, -600 - 50 * i, v=(3 * i, -2 * i, 20))
    # a held pair: the holder's root is what moves (uCenter(him->holdRoot), teleport.c:31)
    h = cube(R, 100, 900, -2500, v=(0, 0, 10))
    k = cube(R, 100, 900, -3400, r=300)
    R.lab_request(h, 2, k, HOLD_MOVE)
    # destination `da` is lit on every tick but 20..24 (pushBtn(da, 2) like a logic gate does, logic.c:76)
    # and `db` is lit on ticks 3..8: pad b teleports, pad c sees two lit destinations and does nothing
    return 60, ([(t, 3, da, -1, 2) for t in range(60) if not 20 <= t < 25] + [(t, 3, db, -1, 2) for t in range(3, 9)])


def lab_bottles(R):
    """bottle2 blocks (PUSH_BOTTLE) and a puddlejumper (PUSH_PUDDLEJUMP) filled by falling bottles"""
    b2 = pad(R, 0, 0)
    R.lab_set_handlers(b2, push=PUSH_BOTTLE)
    R.lab_set_sliders(b2, [0, -1, 0, 100])            # hand, type (unknown), amount, max
    b3 = pad(R, 6000, 0)
    R.lab_set_handlers(b3, push=PUSH_BOTTLE)
    R.lab_set_sliders(b3, [0, 1, 95, 100])            # holds fuel already, nearly full
    pj = pad(R, 12000, 0)
    R.lab_set_handlers(pj, push=PUSH_PUDDLEJUMP)
    R.lab_set_sliders(pj, [10, 60, 0, 0, 0, 0])       # fuel, fuel_max, vel, vel1, grounded, power
    specs = [
        (-500, -300, [0, 0, 30]), (600, 500, [0, 0, 90]), (100, 100, [0, 1, 10]),     # juice, juice, fuel
        (5800, 200, [0, 1, 20]), (6400, -300, [0, 0, 20]),                            # fuel, juice (wrong type)
        (11800, 0, [0, 1, 30]), (12500, 300, [0, 1, 40]), (12000, -600, [0, 0, 40]),  # fuel, fuel, juice
    ]
    for i, (x, y, sl) in enumerate(specs):
        e = cube(R, x, y, -500 - 900 * (i % 3), v=(0, 0, 5 * i), r=250, extra_type=T_EQUIP_SM)
        R.lab_set_handlers(e, tick=TICK_BOTTLE)       # identifies a bottle (bottle.cpp:95); never run (physics-only)
        R.lab_set_sliders(e, sl)
    return 70


def lab_seats(R):
    """seats pick up T_INPUTS cubes that land on them (PUSH_SEAT -> uPickup -> PICKUP_SEAT wires them);
    later the rider is dropped / the seat is killed (FUMBLE_SEAT unwires and zeroes sliders)"""
    seats = []
    for i, fum in enumerate((FUMBLE_SEAT, FUMBLE_SEAT_OLD, FUMBLE_SEAT)):
        s = pad(R, 6000 * i, 0)
        R.lab_set_handlers(s, push=PUSH_SEAT, pair=(fum, -1, PICKUP_SEAT, -1))
        R.lab_set_sliders(s, [5, 6, 7, 8, 9, 10])
        seats.append(s)
    riders = [cube(R, 6000 * i + 100, -50, -700 - 100 * i, v=(0, 0, 30), extra_type=T_INPUTS) for i in range(3)]
    cube(R, 300, 900, -700)                            # no T_INPUTS: just rests on the seat
    script = [(t, 3, riders[2], -1, 0) for t in range(40)]   # holds "get off": seat_push leaves it alone (seat.cpp:12)
    # tick 18: rider 0 asks to be dropped (uDrop -> fumble -> FUMBLE_SEAT on the seat);
    # tick 24: seat 1 dies (its rider is fumbled by killEntNoHandlers, FUMBLE_SEAT_OLD only unwires)
    script += [(18, 0, riders[0], -1, 0), (24, 1, seats[1], -1, 0)]
    return 40, script


def lab_crush(R):
    """entities with crush handlers squeezed between a floor and a ceiling that a falling weight drives"""
    floor = pad(R, 0, 0, r=(6000, 6000, 300), z=400)
    out = []
    for i in range(4):
        e = cube(R, -3000 + 2000 * i, 0, -300, r=400)
        R.lab_set_handlers(e, crush=1 + (i & 1))
        out.append(e)
    lid = R.add_ent((0, 0, -1300), (0, 0, 200), (5000, 5000, 200), GROUND, 0, WM_ME)   # an unstoppable slab
    return 12


TICK_RAND, TICK_LOGIC, TICK_LOGIC_DEBUG, TICK_DOOR, TICK_CURSED, TICK_RESPAWN = 6, 11, 12, 13, 14, 15
TICK_ADJUST, TICK_SEAT, TICK_TIMER, TICK_DEMOLISH, TICK_SIGN = 19, 23, 28, 29, 34


def gate(R, x, y, z, tick, sliders, held_too=True):
    """a small weightless block running a logic-family tick handler"""
    e = R.add_ent((x, y, z), (0, 0, 0), (50, 50, 50), T_WEIGHTLESS, 0, WM_ME)
    R.lab_set_handlers(e, tick=tick, tick_held=tick if held_too else -1)
    R.lab_set_sliders(e, sliders)
    return e


def lab_logic(R):
    """FULL ticks (doUpdates + physics, btref step mode 0): an oscillating timer drives gates, a door
    platform carrying cubes shuttles between two markers, an adjuster retunes it, demolish / cursed
    blocks die, a randomiser picks outputs from gs->rand, a respawner counts down, a seat mirrors its
    rider, and one gate runs as tickHeld.  Scripted pushBtn calls play the part of player input."""
    script = []
    # self-inhibiting timer: outputs, then sees its own pulse on button 0 and waits `delay` ticks
    tm = gate(R, -5000, 0, -3000, TICK_TIMER, [3, 1, 0])
    g_xor = gate(R, -5000, 500, -3000, TICK_LOGIC, [0, 1])        # XOR: fires while the timer pulse is in
    g_xnor = gate(R, -5000, 1000, -3000, TICK_LOGIC_DEBUG, [1, 3])  # XNOR: fires when both or neither
    g_and = gate(R, -5000, 1500, -3000, TICK_LOGIC, [2, 3])       # AND of timer and xnor
    R.lab_add_wire(tm, tm); R.lab_add_wire(tm, g_xor); R.lab_add_wire(tm, g_and)
    R.lab_add_wire(g_xor, g_xnor); R.lab_add_wire(g_xnor, g_and); R.lab_add_wire(g_and, g_xor)
    # door platform with two markers; marker gates are lit by script in turns
    m1 = gate(R, 0, 0, -200, TICK_LOGIC, [0, 0])
    m2 = gate(R, 6000, 2000, -1200, TICK_LOGIC, [0, 0])
    door = R.add_ent((0, 0, 400), (0, 0, 0), (1500, 1500, 300), GROUND, 0, WM_ME)
    R.lab_set_handlers(door, tick=TICK_DOOR)
    R.lab_set_sliders(door, [60, 6])                               # speed, accel
    R.lab_add_wire(door, m1); R.lab_add_wire(door, m2)
    for i, (x, y) in enumerate(((-600, -300), (500, 400), (100, -900))):
        cube(R, x, y, -400 - 30 * i)
    script += [(t, 3, m1, -1, 0) for t in range(0, 25)] + [(t, 3, m2, -1, 0) for t in range(25, 70)]
    script += [(t, 3, m1, -1, 0) for t in range(85, 100)]          # 70..84: nobody lit -> coast to a halt
    # adjuster: while its input is on, retunes the door's speed slider (and a bogus slider index on m2)
    adj = gate(R, -5000, 2000, -3000, TICK_ADJUST, [0, 25])
    R.lab_add_wire(adj, door); R.lab_add_wire(adj, m2)
    script += [(t, 3, adj, -1, 1) for t in range(40, 44)]
    # demolish blocks wired from the AND gate, a cursed block, a sign
    for i in range(2):
        d = cube(R, -9000, 1200 * i, -400, r=300)
        R.lab_set_handlers(d, tick=TICK_DEMOLISH)
        if i == 0:
            R.lab_add_wire(g_and, d)
    floor = pad(R, -9000, 600, r=(2000, 2000, 300), z=400)
    cz = cube(R, -9000, 600, -1500, r=200)
    R.lab_set_handlers(cz, tick=TICK_CURSED)
    sg = gate(R, -5000, 2500, -3000, TICK_SIGN, [1, 6000, 0], held_too=False)
    # randomiser over three gates, two of which are sometimes busy
    rnd = gate(R, -5000, 3000, -3000, TICK_RAND, [1, 3])
    outs = [gate(R, -5000 + 300 * k, 3500, -3000, TICK_LOGIC, [0, 1]) for k in range(3)]
    for o in outs:
        R.lab_add_wire(rnd, o)
    R.lab_add_wire(outs[0], adj)
    script += [(t, 3, outs[1], -1, 0) for t in range(10, 30, 3)] + [(t, 3, rnd, -1, 0) for t in range(50, 55)]
    # respawner with nobody to revive: only its cooldown runs (slider 4)
    rs = gate(R, -5000, 4000, -3000, TICK_RESPAWN, [0, 0, 0, 10, 7, 0, 0])
    script += [(t, 3, rs, -1, 0) for t in range(3, 5)]
    # a seat with a wired rider: mirrors sliders 0-4 and buttons, drops the rider on button 0
    seat = pad(R, 12000, 0)
    R.lab_set_handlers(seat, push=PUSH_SEAT, pair=(FUMBLE_SEAT, -1, PICKUP_SEAT, -1), tick=TICK_SEAT, tick_held=TICK_SEAT)
    R.lab_set_sliders(seat, [0, 0, 0, 0, 0, 0, 0])
    rider = cube(R, 12100, -50, -700, v=(0, 0, 30), extra_type=T_INPUTS)
    R.lab_set_sliders(rider, [11, -12, 13, 14, 15, 16])
    script += [(t, 3, rider, -1, 1 + (t % 3)) for t in range(20, 50)] + [(60, 3, rider, -1, 0)]
    # a gate that is HELD by a cube runs its tickHeld slot
    carrier = cube(R, 18000, 0, -800)
    pad(R, 18000, 0)
    hg = gate(R, 18000, 0, -1400, TICK_LOGIC, [1, 1])
    R.lab_set_handlers(hg, tick=0)                                 # only tickHeld is set: silent until picked up
    lamp = gate(R, 18000, 600, -3000, TICK_DEMOLISH, [])
    R.lab_add_wire(hg, lamp)
    script += [(30, 2, carrier, hg, HOLD_MOVE)]
    return 100, script, 0


REQ_DROP, REQ_DEAD, REQ_PICKUP, REQ_BUTTON, REQ_VEL, REQ_CENTER, REQ_SLIDER, REQ_ERASE, REQ_TYPEMASK, REQ_COLLIDEMASK = range(10)


def lab_requests(R):
    """every kind of host request (bt_world_apply_requests) against the reference's own updaters:
    uVel / uCenter kicks, two uSlider calls on one slider in one tick (min/max averaging), uErase
    (no crush handler call) vs uDead (crush handler call), uMyTypeMask taking an entity out of the
    collision world and back in (it gains a box again: leaf set changes), uMyCollideMask"""
    pad(R, 0, 0, r=(6000, 3000, 300))
    a, b, c, d, e, f, g = [cube(R, -4500 + 1500 * i, 0, -320 - 10 * i) for i in range(7)]
    top = cube(R, -4500 + 1500 * 4, 50, -1300)      # rests on `e`: falls through while e is not collidable
    R.lab_set_sliders(c, [5, 0, -7])
    R.lab_set_handlers(d, crush=1)
    R.lab_set_handlers(g, crush=2)
    script = [
        (3, REQ_VEL, a, -1, 40, -25, -90),
        (5, REQ_CENTER, b, -1, 130, 260, -40),
        (6, REQ_SLIDER, c, -1, 0, 100, 0), (6, REQ_SLIDER, c, -1, 0, -31, 0), (6, REQ_SLIDER, c, -1, 2, 9, 0),
        (6, REQ_SLIDER, c, -1, 7, 1, 0),             # no such slider: ignored (entUpdaters.c:55-60)
        (8, REQ_ERASE, d, -1, 0, 0, 0),
        (10, REQ_TYPEMASK, e, -1, T_DECOR, 0, 0), (10, REQ_COLLIDEMASK, e, -1, 0, 0, 0),
        (12, REQ_COLLIDEMASK, f, -1, T_TERRAIN, 0, 0),
        (14, REQ_DEAD, g, -1, 0, 0, 0),
        (22, REQ_TYPEMASK, e, -1, CUBE, 0, 0), (22, REQ_COLLIDEMASK, e, -1, CUBE_COLLIDE, 0, 0),
        (24, REQ_VEL, top, -1, 0, 0, -150), (24, REQ_VEL, top, -1, 10, 0, 0),   # two uVel in a tick add up
    ]
    return 45, script


def lab_rootorder(R):
    """flushPickups walks gs->rootEnts, newest ROOT first (ent.cpp:775-792), which is not newest ENTITY first once a
    fumble has re-rooted an older entity (addRootEnt prepends, ent.cpp:60-66).  H holds O; O is dropped at tick 3 and
    so becomes the head of the root list, ahead of the younger N; at tick 5 O and N each ask to pick the other up: the
    one visited first ends up HELD (the second request then fails the `x->holdRoot != i` test).  The reference's
    order makes N the holder; a walk in entity order would make O the holder."""
    mk = lambda x: R.add_ent((x, 0, -5000), (0, 0, 0), (50, 50, 50), T_WEIGHTLESS, 0, WM_ME)
    H, O = mk(0), mk(2000)
    R.lab_request(H, 2, O, HOLD_MOVE)       # applied by the first flush1
    filler = [mk(4000 + 2000 * i) for i in range(3)]
    N = mk(12000)
    P, Q = mk(14000), mk(16000)             # the same with the roles of old and young swapped, never fumbled: control
    script = [(3, 0, O, -1, 0),
              (5, 2, N, O, HOLD_MOVE), (5, 2, O, N, HOLD_MOVE),
              (5, 2, P, Q, HOLD_MOVE), (5, 2, Q, P, HOLD_MOVE)]
    return 10, script


# builders return ticks, or (ticks, script[, btref step mode]) with script rows (tick, kind, uid, other, a0[, a1, a2])
# applied before that tick's step; kind = BT_REQ_* (0 uDrop, 1 uDead, 2 uPickup(uid, other, a0), 3 pushBtn(uid, a0), ...)
TICK_SPIDER, TICK_LOADBLOCK, TICK_PUDDLEJUMP, TICK_HELD_GUN, TICK_HELD_BLINK, TICK_HELD_JUMPER = 33, 35, 36, 24, 30, 31
PUSHED_PUDDLEJUMP = 1
REQ_PICKUP, REQ_SLIDER = 2, 6


def lab_pilots(R):
    """FULL ticks: the tick handlers that read a pilot's controls (sliders and buttons of a wired entity: getAxis =
    sliders 0 / 1, player.c:48-51): a puddlejumper steered around over a pad by a scripted input block (its fuel runs
    out on the way), a two-legged spider walking its body across a floor, a loadblock cooling down, and a gun / blink
    / jumper held by a cube WITHOUT inputs (their tickHeld handlers do nothing then).  Scripted uSlider calls on the
    input blocks play the pilot."""
    script = []
    # puddlejumper: falls 16 per tick per tick unless grounded; pushed handler tracks grounded / slide velocity
    pad(R, 0, 0, r=(6000, 6000, 300), z=400)
    pj = R.add_ent((0, 0, -700), (0, 0, 0), (400, 400, 300), CUBE, CUBE_COLLIDE, WM_WOOD)
    R.lab_set_handlers(pj, push=PUSH_PUDDLEJUMP, pushed=PUSHED_PUDDLEJUMP, tick=TICK_PUDDLEJUMP)
    R.lab_set_sliders(pj, [30, 4000, 0, 0, 1, 0])
    stick = gate(R, 0, 0, -5000, 0, [0, 0, 0, 0, 0, 0])
    R.lab_add_wire(pj, stick)
    for t, (ax, ay) in ((3, (64, 0)), (12, (64, -40)), (25, (-10, 64)), (40, (0, 0)), (55, (-64, -64)), (70, (5, 0))):
        script += [(t, REQ_SLIDER, stick, -1, 0, ax, 0), (t, REQ_SLIDER, stick, -1, 1, ay, 0)]
    # a second one without a pilot, airborne: burns its fuel, sinks
    pj2 = R.add_ent((3000, 3000, -4000), (0, 0, 0), (400, 400, 300), CUBE, CUBE_COLLIDE, WM_WOOD)
    R.lab_set_handlers(pj2, push=PUSH_PUDDLEJUMP, pushed=PUSHED_PUDDLEJUMP, tick=TICK_PUDDLEJUMP)
    R.lab_set_sliders(pj2, [12, 4000, 0, 0, 0, 0])
    # spider: a body on a floor holds two legs (tickHeld) wired to one input block
    pad(R, 20000, 0, r=(8000, 8000, 300), z=400)
    body = cube(R, 20000, 0, -1100, r=500)
    walk = gate(R, 20000, 0, -6000, 0, [0, 0])
    legs = []
    for k, dx in enumerate((-700, 700)):
        leg = R.add_ent((20000 + dx, 0, -500), (0, 0, 0), (150, 150, 150), CUBE, CUBE_COLLIDE, WM_WOOD)
        R.lab_set_handlers(leg, tick_held=TICK_SPIDER)
        # d1, d2, scale, t1, t2, t, spring, drag, max accel, ratio
        R.lab_set_sliders(leg, [300, 600, 400, 6, 10, 3 * k, 8, 30, 60, 2])
        R.lab_add_wire(leg, walk)
        legs.append(leg)
        script.append((0, REQ_PICKUP, body, leg, HOLD_MOVE, 0, 0))
    for t, (ax, ay) in ((5, (64, 0)), (30, (30, 50)), (50, (0, 0)), (60, (-64, 20))):
        script += [(t, REQ_SLIDER, walk, -1, 0, ax, 0), (t, REQ_SLIDER, walk, -1, 1, ay, 0)]
    # loadblock cooling down; equipment held by a cube that has no inputs
    lb = gate(R, -8000, 0, -3000, TICK_LOADBLOCK, [3, 9])
    pad(R, -8000, 4000)
    holder = cube(R, -8000, 4000, -800)
    for k, th in enumerate((TICK_HELD_GUN, TICK_HELD_BLINK, TICK_HELD_JUMPER)):
        it = R.add_ent((-8000 + 300 * k, 4000, -1500), (0, 0, 0), (100, 100, 100), CUBE | T_EQUIP_SM, CUBE_COLLIDE, WM_WOOD)
        R.lab_set_handlers(it, tick_held=th)
        R.lab_set_sliders(it, [0, 5, 1, 0, 10, 3, 45, 14000])
        script.append((1, REQ_PICKUP, holder, it, HOLD_MOVE, 0, 0))
    # eyes (eyes.cpp): a head with look sliders holds an eye that tracks where it looks and is dropped later (the eye's own
    # tick handler then makes it debris and retires itself); a vehicle eye follows the pilot's axis through seat wires
    pad(R, -20000, 0, r=(5000, 5000, 300), z=400)
    head = cube(R, -20000, 0, -900, r=500, extra_type=T_INPUTS)
    R.lab_set_sliders(head, [0, 0, 64, 0, 0, 0])
    eye = R.add_ent((-20000, 0, -900), (0, 0, 0), (100, 100, 100), T_DECOR | 128, 0, WM_ME)
    R.lab_set_handlers(eye, pair=(-1, 3, -1, -1), tick=TICK_EYE, tick_held=TICK_HELD_EYE)   # onFumbled = FUMBLED_DECOR
    script.append((0, REQ_PICKUP, head, eye, HOLD_DROP, 0, 0))
    for t, look in ((6, (0, 64, 0)), (15, (40, -40, 20)), (28, (0, 0, -64)), (36, (-64, 10, 3))):
        script += [(t, REQ_SLIDER, head, -1, 2 + k, look[k], 0) for k in range(3)]
    script.append((45, 0, eye, -1, 0, 0, 0))                       # uDrop(eye)
    car = cube(R, -20000, 3000, -900, r=600)
    veye = R.add_ent((-20000, 3000, -900), (0, 0, 0), (100, 100, 100), T_DECOR | 128, 0, WM_ME)
    R.lab_set_handlers(veye, tick_held=TICK_HELD_VEH_EYE)
    vseat = gate(R, -20000, 3000, -7000, 0, [0, 0])
    R.lab_add_wire(veye, vseat)
    R.lab_add_wire(vseat, stick)                                   # the puddlejumper's pilot drives this one too
    script.append((0, REQ_PICKUP, car, veye, HOLD_DROP, 0, 0))
    # leggs (legg.cpp): a walker body on a floor holds two new-style leggs wired to the spider's input block, and one old-style
    # legg wired to the vehicle eye (its input = the eye's offset from the car that holds it)
    pad(R, 20000, 20000, r=(8000, 8000, 300), z=400)
    wbody = cube(R, 20000, 20000, -1100, r=500)
    for k, dx in enumerate((-700, 700)):
        lg = R.add_ent((20000 + dx, 20000, -500), (0, 0, 0), (150, 150, 150), CUBE, CUBE_COLLIDE, WM_WOOD)
        R.lab_set_handlers(lg, tick_held=TICK_LEGG)
        # counter, width, height, stepHeight, tread, speed, accel, time, v_time
        R.lab_set_sliders(lg, [4 * k, 500, 600, 250, 40, 60, 12, 8, 3])
        R.lab_add_wire(lg, walk)
        script.append((0, REQ_PICKUP, wbody, lg, HOLD_MOVE, 0, 0))
    lo = R.add_ent((20000, 20700, -500), (0, 0, 0), (150, 150, 150), CUBE, CUBE_COLLIDE, WM_WOOD)
    R.lab_set_handlers(lo, tick_held=TICK_LEGG_OLD)
    R.lab_set_sliders(lo, [0, 600, 250, 40, 60, 12, 8, 3])   # counter, height, stepHeight, tread, speed, accel, time, v_time
    R.lab_add_wire(lo, veye)
    script.append((0, REQ_PICKUP, wbody, lo, HOLD_MOVE, 0, 0))
    # a player body (player.c) walked, jumped and made to drop / swap its equipment by scripted controls: sliders 0 / 1 = axis,
    # 2..4 = look, 12 = the extra buttons (alt-fire 1 / 2), button 0 = jump.  It never fires: stackem blocks and platforms
    # are entities somebody has to create
    pad(R, -20000, -20000, r=(9000, 9000, 300), z=400)
    pl = R.add_ent((-20000, -20000, -900), (0, 0, 0), (500, 500, 500), T_OBSTACLE | 256 | T_INPUTS, T_OBSTACLE | T_TERRAIN, 5)
    R.lab_set_handlers(pl, push=1, pushed=5, pair=(-1, -1, 5, -1), tick=7)
    R.lab_set_sliders(pl, [0, 0, 0, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    gun = R.add_ent((-20000, -20000, -1500), (0, 0, 0), (120, 120, 120), CUBE | T_EQUIP, CUBE_COLLIDE, WM_WOOD)
    R.lab_set_handlers(gun, tick_held=TICK_HELD_GUN)
    R.lab_set_sliders(gun, [0, 5, 1, 0, 10, 3, 45, 14000])
    charm = R.add_ent((-20300, -20000, -1500), (0, 0, 0), (80, 80, 80), CUBE | T_EQUIP_SM, CUBE_COLLIDE, WM_WOOD)
    R.lab_set_handlers(charm, tick_held=31)      # a jumper (jumper.c): reloads and cools down in the hand, never triggered
    R.lab_set_sliders(charm, [1, 0, 0, 5, 2, 20, 10, 300])
    script += [(0, REQ_PICKUP, pl, gun, HOLD_DROP, 0, 0), (0, REQ_PICKUP, pl, charm, HOLD_DROP, 0, 0)]
    for t, (ax, ay) in ((3, (64, 0)), (20, (0, -64)), (35, (0, 0)), (60, (-40, 40))):
        script += [(t, REQ_SLIDER, pl, -1, 0, ax, 0), (t, REQ_SLIDER, pl, -1, 1, ay, 0)]
    script += [(t, REQ_BUTTON, pl, -1, 0, 0, 0) for t in (10, 11, 12, 50, 51)]          # jump
    script += [(40, REQ_SLIDER, pl, -1, 12, 1, 0), (44, REQ_SLIDER, pl, -1, 12, 0, 0)]      # alt-fire 1: the gun is thrown where it looks
    script += [(55, REQ_SLIDER, pl, -1, 12, 1, 0), (58, REQ_SLIDER, pl, -1, 12, 0, 0)]      # again, hand 1 empty: the charm changes hands
    return 90, script, 0


TICK_LEGG_OLD, TICK_LEGG = 16, 22
TICK_EYE, TICK_HELD_EYE, TICK_HELD_VEH_EYE = 8, 9, 18
TICK_EXPLOSION, TICK_PLATFORM, TICK_THRUST, TICK_BULLET, TICK_HELD_BULLET, TICK_DUST, TICK_INCINERATE = 1, 3, 10, 25, 26, 32, 40
REQ_BUTTON = 3


def lab_ticks2(R):
    """FULL ticks: the short-lived things other handlers leave behind, built by hand (nothing here creates entities): an
    explosion fragment and a dust mote running out, a platform clearing its special behaviour on its first tick, bullets
    in flight timing out and bullets stuck in a block killing it, thrusters pushing a floating body away from their lit
    wires, an incinerator switched on late that kills whatever lies fully beyond its range."""
    script = []
    pad(R, 0, 0, r=(6000, 6000, 300), z=400)
    ex = cube(R, -2000, 0, -800, r=100)
    R.lab_set_handlers(ex, tick=TICK_EXPLOSION); R.lab_set_sliders(ex, [0])
    du = R.add_ent((-1000, 0, -2000), (5, 0, 0), (60, 60, 60), T_WEIGHTLESS, 0, WM_ME)
    R.lab_set_handlers(du, tick=TICK_DUST); R.lab_set_sliders(du, [3, -2, 1, 5])
    pf = R.add_ent((3000, 3000, -1500), (0, 0, 0), (1600, 1600, 250), GROUND, 0, WM_ME)
    R.lab_set_handlers(pf, crush=2, tick=TICK_PLATFORM)
    m = gate(R, 3000, 3000, -4000, 0, [0, 0])
    R.lab_add_wire(pf, m); R.lab_add_wire(pf, ex)
    # bullets in flight
    for k in range(3):
        b = R.add_ent((-3000, 1000 * k, -3000), (40, 0, 0), (100, 100, 100), T_WEIGHTLESS, 0, WM_ME)
        R.lab_set_handlers(b, tick=TICK_BULLET, tick_held=TICK_HELD_BULLET); R.lab_set_sliders(b, [4 + 3 * k])
    # bullets stuck in a plain block (3 kill it), in a debris block (1 kills it) and two in a third block (not enough); they
    # are picked up by the first prepPhysics (ttl 1 -> 0 in the tick before), then ttl runs 0 .. -15
    victim = cube(R, 1500, -1500, -800)
    deb = cube(R, 1500, 1500, -800, extra_type=T_DEBRIS)
    tough = cube(R, 3000, -3000, -800)
    for k, host in enumerate((victim, victim, victim, deb, tough, tough)):
        b = R.add_ent((1500, -1500 + 100 * k, -800), (0, 0, 0), (100, 100, 100), 0, 0, WM_ME)
        R.lab_set_handlers(b, tick=TICK_BULLET, tick_held=TICK_HELD_BULLET); R.lab_set_sliders(b, [1])
        script.append((0, REQ_PICKUP, host, b, HOLD_DROP | HOLD_FREEZE, 0, 0))
    # a floating body with two thrusters wired to lamps; the lamps are lit by script (button 2 = "active")
    body = R.add_ent((-4000, -4000, -3000), (0, 0, 0), (500, 500, 500), CUBE | T_WEIGHTLESS, CUBE_COLLIDE, WM_WOOD)
    lamps = [gate(R, -4000 + 800 * d, -4000, -3000, 0, [0, 0]) for d in (-1, 1)]
    for k in range(2):
        th = R.add_ent((-4000, -4000 + 300 * k, -3600), (0, 0, 0), (100, 100, 100), T_WEIGHTLESS, 0, WM_ME)
        R.lab_set_handlers(th, tick=TICK_THRUST, tick_held=TICK_THRUST); R.lab_set_sliders(th, [640])
        R.lab_add_wire(th, lamps[k])
        script.append((0, REQ_PICKUP, body, th, HOLD_MOVE, 0, 0))
    script += [(t, REQ_BUTTON, lamps[0], -1, 2, 0, 0) for t in range(5, 30)] + [(t, REQ_BUTTON, lamps[1], -1, 2, 0, 0) for t in range(20, 45)]
    # incinerator: off while its two inputs differ (button 0 held by script), then on
    inc = gate(R, 0, 0, -1000, TICK_INCINERATE, [7000])
    script += [(t, REQ_BUTTON, inc, -1, 0, 0, 0) for t in range(0, 40)]
    far = cube(R, 9000, 0, -3000)            # fully beyond 7000 on x: dies when the incinerator comes on
    near = cube(R, 6000, 0, -3000)           # 6000 - 400 < 7000: stays
    # factory.cpp without the spawn (entity creation is the host's): a factory told to finalize its wired product (velocity,
    # masks, buttons, handler slots, then it unwires), and a factory-wires block copying the wires of its inputs to the
    # entity its holder is wired to
    fac = gate(R, -2000, 3000, -3500, 20, [0, 6, -4, 2, CUBE, CUBE_COLLIDE, 0xff00, 2 | 8, WM_WOOD, TICK_BULLET, 0, 0, 0, 2, 0, 3, 0, 0, 25])
    # the product is NEWER than its factory, as after a real spawn() (initEnt prepends to gs->ents): it has had its tick by
    # the time the factory, later in the list, gives it its handlers
    prod = R.add_ent((-2000, 3000, -2500), (0, 0, 0), (200, 200, 200), T_WEIGHTLESS, 0, WM_ME)
    R.lab_set_sliders(prod, [0, 0, 0])
    R.lab_add_wire(fac, prod)
    script.append((5, REQ_BUTTON, fac, -1, 2, 0, 0))
    target = gate(R, -2000, 4000, -3500, 0, [0])
    R.lab_set_handlers(target, pair=(-1, -1, PICKUP_SEAT, -1))   # a seat: its wire block has one spare slot on the device
    hh = cube(R, -2000, 5000, -800)
    R.lab_add_wire(hh, target)
    fw = R.add_ent((-2000, 5000, -1500), (0, 0, 0), (100, 100, 100), T_WEIGHTLESS, 0, WM_ME)
    R.lab_set_handlers(fw, tick=21, tick_held=21)
    srcs = [gate(R, -2500 + 200 * k, 5500, -3500, 0, [0]) for k in range(2)]
    for sgate in srcs:
        R.lab_add_wire(fw, sgate)
    R.lab_add_wire(srcs[0], m); R.lab_add_wire(srcs[1], m)       # the same target twice: one new wire
    script.append((0, REQ_PICKUP, hh, fw, HOLD_MOVE, 0, 0))
    script.append((8, REQ_BUTTON, fw, -1, 0, 0, 0))
    return 60, script, 0


LABS = {
    "lab_pilots": lab_pilots,
    "lab_ticks2": lab_ticks2,
    "lab_teleport": lab_teleport,
    "lab_bottles": lab_bottles,
    "lab_seats": lab_seats,
    "lab_crush": lab_crush,
    "lab_logic": lab_logic,
    "lab_requests": lab_requests,
    "lab_rootorder": lab_rootorder,
}


def make(name, verbose=False):
    from oracle.pybind import RefLib
    from util import golden_path, save_world
    R = RefLib()
    w = R.new_world()
    R.set_gamerules(EFFECT_GRAV)
    res = LABS[name](R)
    script = []
    mode = 1   # btref step mode: 1 = physics only (tick handlers never run), 0 = the full doWholeStep
    if isinstance(res, tuple):
        ticks, script = res[0], res[1]
        if len(res) > 2:
            mode = res[2]
    else:
        ticks = res
    script = [tuple(r) + (0, 0) if len(r) == 5 else tuple(r) for r in script]   # (tick, kind, uid, other, a0, a1, a2)
    R.lab_crush_log()
    wa = R.export()
    hashes = np.zeros(ticks, dtype=np.uint64)
    crush_order = []
    for t in range(ticks):
        for row in script:
            tt, what, uid, other, a0, a1, a2 = row
            if tt == t:
                if what <= 3:
                    R.lab_request(uid, what, other, a0)
                else:
                    R.lab_request3(uid, what, (a0, a1, a2))
        R.step(mode)
        hashes[t] = R.hash()
        for u in R.lab_crush_log():
            crush_order.append((t, u))
    final = R.export()
    n_final = R.count()[0]
    R.free_world(w)
    extra = {"ref_hashes": hashes, "n_final": np.array([n_final]),
             "script": np.array(script, dtype=np.int64).reshape(-1, 7), "mode": np.array([mode]),
             "crush_order": np.array(crush_order, dtype=np.int64).reshape(-1, 2)}
    for k, v in final.a.items():
        extra["final_" + k] = v
    extra["final_n"] = np.array([final.n, int(final.desc.n_sliders), int(final.desc.n_wires)])
    save_world(golden_path(name), wa, **extra)
    print("%-14s ents %3d -> %3d  ticks %3d  last hash %016x  crush handlers run: %s" %
          (name, wa.n, n_final, ticks, int(hashes[-1]), crush_order))
    if verbose:
        a = final.a
        n = final.n
        print("  uid      ", a["uid"][:n].tolist())
        print("  center   ", a["center"].reshape(-1, 3)[:n].tolist())
        print("  holder   ", a["holder"][:n].tolist())
        so = a["slider_off"]
        print("  sliders  ", [a["slider_v"][so[i]:so[i + 1]].tolist() for i in range(n)])
        wo = a["wire_off"]
        print("  wires    ", [a["wire_to"][wo[i]:wo[i + 1]].tolist() for i in range(n)])
        print("  btn      ", a["btn"].reshape(-1, 8)[:n].tolist())
        print("  distinct hashes", len(set(hashes.tolist())))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    for nm in (args or list(LABS)):
        make(nm, verbose="-v" in sys.argv)
