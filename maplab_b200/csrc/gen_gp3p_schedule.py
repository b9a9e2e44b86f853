#!/usr/bin/env python3
"""Re-schedule the GP3P elimination micro-op program (gp3p_program.inc) for warp-parallel execution.

The sequential program (12 253 micro-ops over 1 569 slots, derived from opengv's generated GP3P
Groebner template by oracle/gen_gp3p_program.py) has a critical path of only a few hundred ops.
This script
  1. makes the implicit `factor` register explicit (three-address form with factor slots),
  2. removes dead operations (results that never reach the 48 action-matrix entries),
  3. builds the true (read-after-write) dependency graph of the remaining operations,
  4. list-schedules them (longest-path priority) into STEPS of at most 32 mutually independent
     operations — one step = one warp-wide instruction group,
  5. allocates a physical slot per value by liveness over the steps (in-place updates where
     possible), which removes the WAR/WAW hazards of the original slot names and shrinks the
     per-hypothesis slot array ~3x (more hypotheses resident per SM),
  6. sorts every step by opcode, pads it to 32 operations and emits gp3p_schedule.inc.
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
(W_DIVSUB, W_DIV, W_NEGDIV, W_ZERO, W_SUBMUL, W_SCALE, W_COPY, W_INV, W_NOP) = range(9)
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

    # ---- 3. values (SSA over the program order) and true dependencies ----
    W_RMW = (W_SUBMUL, W_SCALE)
    cur, vals = {}, []            # logical slot -> value id; value = [kind, producer op, in-place input]
    def new_value(kind, prod, inplace=None):
        vals.append([kind, prod, inplace])
        return len(vals) - 1
    init_value = {}
    for e in init:
        init_value[e[0]] = cur[e[0]] = new_value("init", None)
    op_reads, op_write = [], []
    for i, op in enumerate(ops):
        r, d = reads_writes(op)
        ids = []
        for sl in r:
            if sl not in cur:
                cur[sl] = new_value("zero", None)   # implicit initial 0.0 (structural zero)
            ids.append(cur[sl])
        op_reads.append(ids)
        cur[d] = v = new_value("def", i, ids[0] if op[0] in W_RMW else None)
        op_write.append(v)
    out_value = {sl: cur[sl] for sl in action if sl >= 0}
    n = len(ops)
    succs, npred = [[] for _ in range(n)], [0] * n
    for i in range(n):
        preds = set(vals[v][1] for v in op_reads[i] if vals[v][1] is not None)
        npred[i] = len(preds)
        for q in preds:
            succs[q].append(i)
    height = [1] * n              # longest dependency chain from the op to a sink (priority)
    for i in range(n - 1, -1, -1):
        for q in succs[i]:
            height[i] = max(height[i], height[q] + 1)

    # ---- 4. list scheduling into STEPS of at most 32 mutually independent operations ----
    # (one step = one warp-wide instruction group; only true dependencies constrain the order,
    # the slot allocator below removes the WAR/WAW hazards of the original slot names)
    import heapq
    LANES = 32
    ready = [(-height[i], i) for i in range(n) if npred[i] == 0]
    heapq.heapify(ready)
    steps, step_of, left = [], [0] * n, n
    while left:
        take = []
        while ready and len(take) < LANES:
            take.append(heapq.heappop(ready)[1])
        assert take
        for i in take:
            step_of[i] = len(steps) + 1          # steps are numbered from 1; 0 = initial state
        steps.append(take)
        left -= len(take)
        for i in take:
            for q in succs[i]:
                npred[q] -= 1
                if npred[q] == 0:
                    heapq.heappush(ready, (-height[q], q))
    num_waves = len(steps)
    critical_path = max(height)

    # ---- 5. slot allocation by liveness over the steps ----
    # A value lives from the step that defines it to the last step that reads it. Its slot is
    # recycled for definitions of strictly later steps; SUBMUL / SCALE overwrite their first
    # operand in place when they are its only remaining reader. Values that are only ever the
    # implicit initial 0.0 share slot 0.
    END = num_waves + 1
    def_step = [0] * len(vals)
    last_read = [0] * len(vals)
    readers_in_step = {}
    for i in range(n):
        def_step[op_write[i]] = step_of[i]
        last_read[op_write[i]] = max(last_read[op_write[i]], step_of[i])
        for v in set(op_reads[i]):
            last_read[v] = max(last_read[v], step_of[i])
            readers_in_step[(v, step_of[i])] = readers_in_step.get((v, step_of[i]), 0) + 1
    for v in out_value.values():
        last_read[v] = END
    ZERO_SLOT = 0
    phys = [None] * len(vals)
    next_slot, free, expiring, taken_over = 1, [], {}, set()
    for v, (kind, prod, inplace) in enumerate(vals):
        if kind == "zero":
            phys[v] = ZERO_SLOT
        elif kind == "init":
            phys[v] = next_slot
            next_slot += 1
            expiring.setdefault(last_read[v], []).append(v)
    free.extend(phys[v] for v in expiring.pop(0, []))   # init values nobody reads
    for t in range(1, END):
        for i in steps[t - 1]:
            v = op_write[i]
            src = vals[v][2]
            if (src is not None and vals[src][0] != "zero" and last_read[src] == t
                    and readers_in_step[(src, t)] == 1):
                phys[v] = phys[src]                   # in place
                taken_over.add(src)
            elif free:
                phys[v] = free.pop()
            else:
                phys[v] = next_slot
                next_slot += 1
            expiring.setdefault(last_read[v], []).append(v)
        free.extend(phys[v] for v in expiring.pop(t, []) if v not in taken_over)
    compact_slots = next_slot
    new_waves = []
    for take in steps:
        row = []
        for i in take:
            o = ops[i][0]
            rd = [phys[v] for v in op_reads[i]]
            d = phys[op_write[i]]
            if o == W_DIVSUB:
                row.append((o, d, rd[0], rd[1], rd[2], rd[3]))
            elif o in (W_DIV, W_NEGDIV):
                row.append((o, d, rd[0], rd[1], 0, 0))
            elif o == W_ZERO:
                row.append((o, d, 0, 0, 0, 0))
            elif o == W_SUBMUL:                       # d = x - a * b with x the first operand
                row.append((o, d, rd[1], rd[2], rd[0], 0))
            elif o == W_SCALE:                        # d = a * x
                row.append((o, d, rd[1], rd[0], 0, 0))
            else:
                row.append((o, d, rd[0], 0, 0, 0))
        row.sort(key=lambda op: (op[0], op[1]))
        row += [(W_NOP, 0, 0, 0, 0, 0)] * (LANES - len(row))   # padding: no memory access at all
        new_waves.append(row)
    new_init = []
    for e in init:
        v = init_value[e[0]]
        dead = last_read[v] == 0
        new_init.append([-1 if dead else phys[v]] + list(e[1:]))   # -1: nobody reads it, not stored
    new_action = [(-1 if sl < 0 else phys[out_value[sl]]) for sl in action]
    old_total_slots, old_init = total_slots, init
    total_slots, waves = compact_slots, new_waves
    flat = [op for w in waves for op in w]
    offsets = [LANES * i for i in range(num_waves + 1)]
    chunks = num_waves

    # ---- 6. self-check against the sequential program ----
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
            if e[0] >= 0:
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
                if o == W_NOP:
                    continue
                if o == W_DIVSUB:
                    val = S[a] / S[b] - S[c] / S[e]
                elif o == W_DIV:
                    val = S[a] / S[b]
                elif o == W_NEGDIV:
                    val = -S[a] / S[b]
                elif o == W_ZERO:
                    val = 0.0
                elif o == W_SUBMUL:
                    val = S[c] - S[a] * S[b]
                elif o == W_SCALE:
                    val = S[a] * S[b]
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
        fo.write("// op: 0 DIVSUB d=a/b-c/e, 1 DIV d=a/b, 2 NEGDIV d=-a/b, 3 ZERO, 4 SUBMUL d=c-a*b, 5 SCALE d=a*b, 6 COPY d=a, 7 INV d=1/a, 8 NOP\n")
        fo.write("// every step holds exactly 32 operations (padded with NOP): step s = ops [32 s, 32 s + 32); init slot -1 = dead\n")
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
    print(f"ops {len(mops)} -> {n} after DCE; critical path {critical_path}; {num_waves} steps of 32 lanes "
          f"({100.0 * n / (32 * num_waves):.0f}% filled); slots {old_total_slots} -> {total_slots}")


if __name__ == "__main__":
    main()
