#!/usr/bin/env python3
"""Re-schedule the GP3P elimination micro-op program (gp3p_program.inc) for warp-parallel execution.

The sequential program (12 253 micro-ops over 1 569 slots, derived from opengv's generated GP3P
Groebner template by oracle/gen_gp3p_program.py) has a critical path of only a few hundred ops.
This script
  1. makes the implicit `factor` register explicit (three-address form with factor slots),
  2. removes dead operations (results that never reach the 48 action-matrix entries),
  3. list-schedules the rest into WAVES of mutually independent operations (RAW, WAR and WAW all
     respected, so the operations of one wave may execute in any order / in parallel),
  4. renames every value to a recycled physical slot (liveness over the wave schedule), which
     shrinks the per-hypothesis slot array ~3x (more hypotheses resident per SM),
  5. sorts every wave by opcode and emits gp3p_schedule.inc.
Every operation still computes exactly the same IEEE-754 expression on exactly the same operand
values, so the wave-parallel execution is bit-identical to the sequential program; the script
checks that on random inputs before writing the file.

Usage: python maplab_b200/csrc/gen_gp3p_schedule.py
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gp3p_program.inc")
OUT = os.path.join(HERE, "gp3p_schedule.inc")

(MOP_DIVSUB, MOP_DIV, MOP_NEGDIV, MOP_FACTOR_DIV, MOP_ZERO, MOP_SUBMUL, MOP_FACTOR_LOAD,
 MOP_FACTOR_INV, MOP_SCALE) = range(9)
# three-address opcodes
(W_DIVSUB, W_DIV, W_NEGDIV, W_ZERO, W_SUBMUL, W_SCALE, W_COPY, W_INV) = range(8)
FACTOR_POOL = 96


def parse(name, text):
    m = re.search(name + r"\[[^\]]*\](?:\[[^\]]*\])? = \{(.*?)\};", text, re.S)
    rows = re.findall(r"\{([^{}]*)\}", m.group(1))
    if rows:
        return [list(map(int, r.split(","))) for r in rows]
    return list(map(int, m.group(1).split(",")))


def main():
    text = open(SRC).read()
    num_slots = int(re.search(r"#define GP3P_NUM_SLOTS (\d+)", text).group(1))
    init, mops, action = parse("GP3P_INIT", text), parse("GP3P_MOPS", text), parse("GP3P_ACTION", text)

    # ---- 1. three-address form: (op, dst, srcs...) with explicit factor slots ----
    ops = []  # (op, d, a, b, c, e)
    factor_slot, next_factor = None, 0
    for m in mops:
        o = m[0]
        if o == MOP_DIVSUB:
            ops.append((W_DIVSUB, m[1], m[2], m[3], m[4], m[5]))
        elif o == MOP_DIV:
            ops.append((W_DIV, m[1], m[2], m[3], 0, 0))
        elif o == MOP_NEGDIV:
            ops.append((W_NEGDIV, m[1], m[2], m[3], 0, 0))
        elif o == MOP_ZERO:
            ops.append((W_ZERO, m[1], 0, 0, 0, 0))
        elif o in (MOP_FACTOR_DIV, MOP_FACTOR_LOAD, MOP_FACTOR_INV):
            factor_slot = num_slots + (next_factor % FACTOR_POOL)
            next_factor += 1
            if o == MOP_FACTOR_DIV:
                ops.append((W_DIV, factor_slot, m[1], m[2], 0, 0))
            elif o == MOP_FACTOR_LOAD:
                ops.append((W_COPY, factor_slot, m[1], 0, 0, 0))
            else:
                ops.append((W_INV, factor_slot, m[1], 0, 0, 0))
        elif o == MOP_SUBMUL:
            ops.append((W_SUBMUL, m[1], factor_slot, m[2], 0, 0))   # d = d - f * a
        elif o == MOP_SCALE:
            ops.append((W_SCALE, m[1], factor_slot, 0, 0, 0))       # d = f * d
        else:
            raise ValueError(o)
    total_slots = num_slots + FACTOR_POOL

    def reads_writes(op):
        o, d, a, b, c, e = op
        if o == W_DIVSUB:
            return [a, b, c, e], d
        if o in (W_DIV, W_NEGDIV):
            return [a, b], d
        if o == W_ZERO:
            return [], d
        if o == W_SUBMUL:
            return [d, a, b], d
        if o == W_SCALE:
            return [d, a], d
        if o in (W_COPY, W_INV):
            return [a], d
        raise ValueError(o)

    # ---- 2. dead-code elimination (backwards liveness) ----
    live = set(s for s in action if s >= 0)
    keep = [False] * len(ops)
    for i in range(len(ops) - 1, -1, -1):
        r, w = reads_writes(ops[i])
        if w in live:
            keep[i] = True
            if ops[i][0] not in (W_SUBMUL, W_SCALE):
                live.discard(w)  # a pure definition kills the slot
            live.update(r)
    ops = [op for op, k in zip(ops, keep) if k]

    # ---- 3. wave scheduling ----
    last_write = {}   # slot -> level of its last writer
    last_read = {}    # slot -> max level of readers since the last write
    level = []
    for op in ops:
        r, w = reads_writes(op)
        lv = 0
        for s in r:
            lv = max(lv, last_write.get(s, 0) + 1)          # RAW
        lv = max(lv, last_write.get(w, 0) + 1)              # WAW
        lv = max(lv, last_read.get(w, 0) + 1)               # WAR
        lv = max(lv, 1)
        level.append(lv)
        for s in r:
            last_read[s] = max(last_read.get(s, 0), lv)
        last_write[w] = lv
        last_read[w] = 0 if w not in r else lv
    num_waves = max(level)
    waves = [[] for _ in range(num_waves)]
    for op, lv in zip(ops, level):
        waves[lv - 1].append(op)
    for w in waves:
        w.sort(key=lambda op: (op[0], op[1]))
    flat = [op for w in waves for op in w]
    offsets = np.cumsum([0] + [len(w) for w in waves]).tolist()
    chunks = sum((len(w) + 31) // 32 for w in waves)

    # ---- 4. slot compaction: rename every VALUE (definition) to a physical slot that is free ----
    # A value lives from the wave that defines it (0 = initial state) to the last wave that reads it;
    # SUBMUL / SCALE update their destination in place and keep its slot. A slot is recycled for a
    # definition of wave w only if its previous value was last read in a wave < w, so the
    # operations of one wave stay mutually independent. Slots that are only ever read as the
    # implicit initial 0.0 (structural zeros of the elimination template) share slot 0.
    W_RMW = (W_SUBMUL, W_SCALE)
    cur, vals = {}, []            # logical slot -> value id; value = [def_wave, last_read, kind]
    def new_value(wave, kind):
        vals.append([wave, wave, kind])
        return len(vals) - 1
    init_value = {}
    for e in init:
        init_value[e[0]] = cur[e[0]] = new_value(0, "init")
    op_vals = []                  # per op: (value ids of the reads, value id of the result)
    for wi, w in enumerate(waves):
        reads_of = []
        for op in w:
            r, d = reads_writes(op)
            ids = []
            for sl in r:
                if sl not in cur:
                    cur[sl] = new_value(0, "zero")
                v = cur[sl]
                vals[v][1] = max(vals[v][1], wi + 1)
                ids.append(v)
            reads_of.append(ids)
        for op, ids in zip(w, reads_of):
            d = op[1]
            if op[0] in W_RMW:
                v = cur[d]
                if vals[v][2] == "zero":
                    vals[v][2] = "zero_rmw"   # needs its own zero-filled slot
            else:
                cur[d] = v = new_value(wi + 1, "def")
            op_vals.append((ids, v))
    END = len(waves) + 1
    for sl in action:
        if sl >= 0:
            vals[cur[sl]][1] = END
    ZERO_SLOT = 0
    phys = [None] * len(vals)
    next_slot, free = 1, []
    by_def = {}
    for v, (dw, lr, kind) in enumerate(vals):
        by_def.setdefault(dw, []).append(v)
    expiring = {}
    for wave in range(0, END + 1):
        for v in by_def.get(wave, []):
            dw, lr, kind = vals[v]
            if kind == "zero":
                phys[v] = ZERO_SLOT
                continue
            # initial values (wave 0) need fresh slots: they are written / zero-filled before wave 1
            if wave > 0 and free:
                phys[v] = free.pop()
            else:
                phys[v] = next_slot
                next_slot += 1
            expiring.setdefault(lr, []).append(phys[v])
        # values last read in this wave free their slot for definitions of LATER waves
        free.extend(expiring.pop(wave, []))
    dummy_slot = next_slot        # init entries nobody reads land here
    compact_slots = next_slot + 1
    new_waves, at = [], 0
    for w in waves:
        nw_ = []
        for op in w:
            ids, v = op_vals[at]
            at += 1
            r, d = reads_writes(op)
            m = {sl: phys[i] for sl, i in zip(r, ids)}
            o, d0, a0, b0, c0, e0 = op
            def mp(x, used):
                return m[x] if used else 0
            if o == W_DIVSUB:
                nw_.append((o, phys[v], m[a0], m[b0], m[c0], m[e0]))
            elif o in (W_DIV, W_NEGDIV):
                nw_.append((o, phys[v], m[a0], m[b0], 0, 0))
            elif o == W_ZERO:
                nw_.append((o, phys[v], 0, 0, 0, 0))
            elif o == W_SUBMUL:
                assert phys[v] == m[d0]
                nw_.append((o, phys[v], m[a0], m[b0], 0, 0))
            elif o == W_SCALE:
                assert phys[v] == m[d0]
                nw_.append((o, phys[v], m[a0], 0, 0, 0))
            else:
                nw_.append((o, phys[v], m[a0], 0, 0, 0))
        nw_.sort(key=lambda op: (op[0], op[1]))
        new_waves.append(nw_)
    new_init = []
    for e in init:
        v = init_value[e[0]]
        dead = vals[v][1] == 0
        new_init.append([dummy_slot if dead else phys[v]] + list(e[1:]))
    new_action = [(-1 if sl < 0 else phys[cur[sl]]) for sl in action]
    old_total_slots, old_init, old_waves = total_slots, init, waves
    total_slots, waves = compact_slots, new_waves
    flat = [op for w in waves for op in w]

    # ---- 5. self-check against the sequential program ----
    rng = np.random.default_rng(0)

    def init_slots_seq(f, v, p):
        S = np.zeros(old_total_slots)
        src = (f, v, p)
        for e in old_init:
            acc = 0.0
            for t in range(e[1]):
                coef, kind, i, j = e[2 + 4 * t: 6 + 4 * t]
                term = float(coef) * src[kind][j * 3 + i]
                acc = term if t == 0 else acc + term
            S[e[0]] = acc
        return S

    init = new_init

    def init_slots(f, v, p):
        S = np.zeros(total_slots)
        src = (f, v, p)
        for e in init:
            acc = 0.0
            for t in range(e[1]):
                coef, kind, i, j = e[2 + 4 * t: 6 + 4 * t]
                term = float(coef) * src[kind][j * 3 + i]
                acc = term if t == 0 else acc + term
            S[e[0]] = acc
        return S

    def run_seq(S):
        factor = 0.0
        for m in mops:
            o = m[0]
            if o == MOP_DIVSUB:
                S[m[1]] = S[m[2]] / S[m[3]] - S[m[4]] / S[m[5]]
            elif o == MOP_DIV:
                S[m[1]] = S[m[2]] / S[m[3]]
            elif o == MOP_NEGDIV:
                S[m[1]] = -S[m[2]] / S[m[3]]
            elif o == MOP_FACTOR_DIV:
                factor = S[m[1]] / S[m[2]]
            elif o == MOP_ZERO:
                S[m[1]] = 0.0
            elif o == MOP_SUBMUL:
                S[m[1]] = S[m[1]] - factor * S[m[2]]
            elif o == MOP_FACTOR_LOAD:
                factor = S[m[1]]
            elif o == MOP_FACTOR_INV:
                factor = 1.0 / S[m[1]]
            elif o == MOP_SCALE:
                S[m[1]] = factor * S[m[1]]
        return S

    def run_waves(S):
        for w in waves:
            res = []
            for (o, d, a, b, c, e) in reversed(w):  # any order inside a wave; reads before writes
                if o == W_DIVSUB:
                    val = S[a] / S[b] - S[c] / S[e]
                elif o == W_DIV:
                    val = S[a] / S[b]
                elif o == W_NEGDIV:
                    val = -S[a] / S[b]
                elif o == W_ZERO:
                    val = 0.0
                elif o == W_SUBMUL:
                    val = S[d] - S[a] * S[b]
                elif o == W_SCALE:
                    val = S[a] * S[d]
                elif o == W_COPY:
                    val = S[a]
                elif o == W_INV:
                    val = 1.0 / S[a]
                res.append((d, val))
            for d, val in res:
                S[d] = val
        return S

    with np.errstate(all="ignore"):
        for trial in range(5):
            f, v, p = rng.normal(size=9), rng.normal(size=9) * 0.1, rng.normal(size=9) * 3
            a = run_seq(init_slots_seq(f, v, p))
            b = run_waves(init_slots(f, v, p))
            assert b[ZERO_SLOT] == 0.0
            for s, ns in zip(action, new_action):
                if s >= 0:
                    assert a[s].tobytes() == b[ns].tobytes(), (trial, s, a[s], b[ns])

    with open(OUT, "w") as fo:
        fo.write("// GENERATED by maplab_b200/csrc/gen_gp3p_schedule.py from gp3p_program.inc. Do not edit.\n")
        fo.write("// Wave-scheduled three-address form of the GP3P elimination: the operations of one wave are\n")
        fo.write("// mutually independent (RAW/WAR/WAW respected), results are bit-identical to the sequential program.\n")
        fo.write(f"#define GP3P_W_NUM_SLOTS {total_slots}\n#define GP3P_W_NUM_OPS {len(flat)}\n")
        fo.write(f"#define GP3P_W_NUM_WAVES {num_waves}\n#define GP3P_W_NUM_CHUNKS {chunks}\n")
        fo.write("// op: 0 DIVSUB d=a/b-c/e, 1 DIV d=a/b, 2 NEGDIV d=-a/b, 3 ZERO, 4 SUBMUL d=d-a*b, 5 SCALE d=a*d, 6 COPY d=a, 7 INV d=1/a\n")
        fo.write("static const unsigned short GP3P_W_OPS[GP3P_W_NUM_OPS][6] = {\n")
        for op in flat:
            fo.write("  {" + ",".join(str(x) for x in op) + "},\n")
        fo.write("};\nstatic const unsigned short GP3P_W_WAVE_OFFSETS[GP3P_W_NUM_WAVES + 1] = {")
        fo.write(",".join(str(x) for x in offsets) + "};\n")
        fo.write("// GP3P_INIT / GP3P_ACTION of gp3p_program.inc with the compacted slot numbers\n")
        fo.write(f"static const short GP3P_W_INIT[{len(new_init)}][{len(new_init[0])}] = {{\n")
        for e in new_init:
            fo.write("  {" + ",".join(str(x) for x in e) + "},\n")
        fo.write("};\n")
        fo.write(f"static const short GP3P_W_ACTION[{len(new_action)}] = {{" + ",".join(str(x) for x in new_action) + "};\n")
    print(f"ops {len(mops)} -> {len(flat)} after DCE; waves {num_waves}; 32-wide chunks {chunks}; "
          f"slots {old_total_slots} -> {total_slots} after compaction; widest wave {max(len(w) for w in waves)}")


if __name__ == "__main__":
    main()
