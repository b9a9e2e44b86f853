#!/usr/bin/env python3
"""Derive the GP3P Groebner-elimination *program table* from the reference.

opengv's generalised-P3P solver (A20) is machine-generated
straight-line code over a 48x85 double matrix:
  dependencies/3rdparty/opengv/src/absolute_pose/modules/gp3p/init.cpp:33-186
  dependencies/3rdparty/opengv/src/absolute_pose/modules/gp3p/code.cpp:33-913
  dependencies/3rdparty/opengv/src/absolute_pose/modules/gp3p/reductors.cpp
  dependencies/3rdparty/opengv/src/absolute_pose/modules/gp3p/spolynomials.cpp
The elimination template (which entry is combined with which, in which order)
IS the algorithm. This script parses those files where they lie under
/root/reference (dev container only) and re-expresses the template as a flat
table of micro-operations over a dense numbering of the structurally non-zero
matrix entries ("slots"). The same table drives the CPU oracle
(oracle/pnp.cc) and the CUDA kernel (maplab_b200/csrc/ransac_kernels.cu), so
both execute the identical IEEE-754 operation sequence.

Self-check: the parsed statements are also executed literally (dense 48x85
numpy matrix) on random inputs and compared bit-for-bit against the micro-op
interpretation.

Usage: python oracle/gen_gp3p_program.py   (writes both .inc files)
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/dependencies/3rdparty/opengv/src/absolute_pose/modules/gp3p"
HERE = os.path.dirname(os.path.abspath(__file__))
OUTS = [os.path.join(HERE, "gp3p_program.inc"),
        os.path.join(HERE, "..", "maplab_b200", "csrc", "gp3p_program.inc")]

# micro-op codes
(MOP_DIVSUB, MOP_DIV, MOP_NEGDIV, MOP_FACTOR_DIV, MOP_ZERO, MOP_SUBMUL,
 MOP_FACTOR_LOAD, MOP_FACTOR_INV, MOP_SCALE) = range(9)

G = r"groebnerMatrix\((\w+),(\d+)\)"


def parse_init():
    """-> list of (row, col, [(coef, kind, i, j), ...]) ; kind 0=f 1=v 2=p"""
    out = []
    kinds = {"f": 0, "v": 1, "p": 2}
    for line in open(os.path.join(REF, "init.cpp")):
        m = re.match(r"\s*groebnerMatrix\((\d+),(\d+)\) = (.*);", line)
        if not m:
            continue
        row, col, expr = int(m.group(1)), int(m.group(2)), m.group(3)
        expr = expr.replace("(", " ").replace(")", " ")
        # tokens like: -1*f 0,1   | 2*p 1,0 -2*p 1,1 | v 0,0 -v 0,1 +p 0,0 -p 0,1
        toks = re.findall(r"([+-]?)\s*(?:(\d+)\*)?([fvp])\s+(\d),(\d)", expr)
        assert toks, line
        terms = []
        for sign, mul, kind, i, j in toks:
            coef = int(mul) if mul else 1
            if sign == "-":
                coef = -coef
            terms.append((coef, kinds[kind], int(i), int(j)))
        # make sure we consumed the whole expression
        rebuilt = re.sub(r"[\s+\-*\d,fvp]", "", expr)
        assert rebuilt == "", (line, rebuilt)
        out.append((row, col, terms))
    return out


def parse_spolys():
    """name -> list of (row, col, kind, a_row, a_col, a_lead, b_row, b_col, b_lead)"""
    spolys = {}
    cur = None
    for line in open(os.path.join(REF, "spolynomials.cpp")):
        m = re.search(r"gp3p::(sPolynomial\d+)\(", line)
        if m:
            cur = m.group(1)
            spolys[cur] = []
            continue
        line = line.strip()
        m = re.match(G + r" = \(" + G + r"/\(" + G + r"\)-" + G + r"/\(" + G + r"\)\);", line)
        if m:
            g = m.groups()
            r, c = int(g[0]), int(g[1])
            assert g[2] == g[4] and g[6] == g[8]
            spolys[cur].append((r, c, MOP_DIVSUB, int(g[2]), int(g[3]), int(g[5]),
                                int(g[6]), int(g[7]), int(g[9])))
            continue
        m = re.match(G + r" = -" + G + r"/\(" + G + r"\);", line)
        if m:
            g = m.groups()
            assert g[2] == g[4]
            spolys[cur].append((int(g[0]), int(g[1]), MOP_NEGDIV, 0, 0, 0,
                                int(g[2]), int(g[3]), int(g[5])))
            continue
        m = re.match(G + r" = " + G + r"/\(" + G + r"\);", line)
        if m:
            g = m.groups()
            assert g[2] == g[4]
            spolys[cur].append((int(g[0]), int(g[1]), MOP_DIV, int(g[2]), int(g[3]), int(g[5]),
                                0, 0, 0))
            continue
        assert "groebnerMatrix" not in line or "Eigen::Matrix" in line, line
    return spolys


def parse_reductors():
    """name -> (src_row, lead_tcol, lead_scol, [(tcol, scol), ...])"""
    red = {}
    cur = None
    for line in open(os.path.join(REF, "reductors.cpp")):
        m = re.search(r"gp3p::(groebnerRow\w+)\(", line)
        if m:
            cur = m.group(1)
            red[cur] = None
            continue
        line = line.strip()
        m = re.match(r"double factor = groebnerMatrix\(targetRow,(\d+)\) / groebnerMatrix\((\d+),(\d+)\);", line)
        if m:
            red[cur] = [int(m.group(2)), int(m.group(1)), int(m.group(3)), []]
            continue
        m = re.match(r"groebnerMatrix\(targetRow,(\d+)\) = 0.0;", line)
        if m:
            assert int(m.group(1)) == red[cur][1]
            continue
        m = re.match(r"groebnerMatrix\(targetRow,(\d+)\) -= factor \* groebnerMatrix\((\d+),(\d+)\);", line)
        if m:
            assert int(m.group(2)) == red[cur][0]
            red[cur][3].append((int(m.group(1)), int(m.group(3))))
            continue
        assert "groebnerMatrix" not in line or "Eigen::Matrix" in line, line
    return red


def parse_code():
    """-> list of high-level ops"""
    ops = []
    lines = [l.strip() for l in open(os.path.join(REF, "code.cpp"))]
    i = 0
    pending_factor = None
    for line in lines:
        m = re.match(r"(sPolynomial\d+)\(groebnerMatrix\);", line)
        if m:
            ops.append(("spoly", m.group(1)))
            continue
        m = re.match(r"(groebnerRow\w+)\(groebnerMatrix,(\d+)\);", line)
        if m:
            ops.append(("reduce", m.group(1), int(m.group(2))))
            continue
        m = re.match(r"factor = groebnerMatrix\((\d+),(\d+)\);", line)
        if m:
            pending_factor = ("load", int(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"factor = 1.0 / groebnerMatrix\((\d+),(\d+)\);", line)
        if m:
            pending_factor = ("inv", int(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"groebnerMatrix.row\((\d+)\) = groebnerMatrix.row\((\d+)\) - factor \* groebnerMatrix.row\((\d+)\);", line)
        if m:
            t, t2, s = int(m.group(1)), int(m.group(2)), int(m.group(3))
            assert t == t2 and pending_factor[0] == "load"
            ops.append(("rowsub", t, s, pending_factor[1], pending_factor[2]))
            pending_factor = None
            continue
        m = re.match(r"groebnerMatrix.row\((\d+)\) = factor \* groebnerMatrix.row\((\d+)\);", line)
        if m:
            t, t2 = int(m.group(1)), int(m.group(2))
            assert t == t2 and pending_factor[0] == "inv"
            ops.append(("rowscale", t, pending_factor[1], pending_factor[2]))
            pending_factor = None
            continue
        assert "groebnerMatrix" not in line or "Eigen::Matrix" in line, line
    return ops


def dense_run(init, spolys, red, code, f, v, p):
    """Literal dense execution (numpy float64 scalars = IEEE double, no FMA)."""
    M = np.zeros((48, 85), dtype=np.float64)
    src = (f, v, p)
    for row, col, terms in init:
        acc = None
        for coef, kind, i, j in terms:
            t = np.float64(coef) * src[kind][i, j]
            acc = t if acc is None else acc + t
        M[row, col] = acc
    for op in code:
        if op[0] == "spoly":
            for (r, c, kind, a, ca, la, b, cb, lb) in spolys[op[1]]:
                if kind == MOP_DIVSUB:
                    M[r, c] = M[a, ca] / M[a, la] - M[b, cb] / M[b, lb]
                elif kind == MOP_DIV:
                    M[r, c] = M[a, ca] / M[a, la]
                else:
                    M[r, c] = -M[b, cb] / M[b, lb]
        elif op[0] == "reduce":
            s, lt, ls, pairs = red[op[1]]
            t = op[2]
            factor = M[t, lt] / M[s, ls]
            M[t, lt] = 0.0
            for tc, sc in pairs:
                M[t, tc] -= factor * M[s, sc]
        elif op[0] == "rowsub":
            _, t, s, fr, fc = op
            factor = M[fr, fc]
            M[t, :] = M[t, :] - factor * M[s, :]
        elif op[0] == "rowscale":
            _, t, fr, fc = op
            factor = 1.0 / M[fr, fc]
            M[t, :] = factor * M[t, :]
    return M


def build_microops(init, spolys, red, code):
    """Symbolic sparsity propagation + flattening to micro-ops over slots."""
    pattern = [set() for _ in range(48)]
    for row, col, _ in init:
        pattern[row].add(col)
    slot_of = {}

    def slot(r, c):
        key = (r, c)
        if key not in slot_of:
            slot_of[key] = len(slot_of)
        return slot_of[key]

    # Init entries get the first slots, in file order.
    for row, col, _ in init:
        slot(row, col)
    mops = []
    for op in code:
        if op[0] == "spoly":
            for (r, c, kind, a, ca, la, b, cb, lb) in spolys[op[1]]:
                pattern[r].add(c)
                if kind == MOP_DIVSUB:
                    mops.append((MOP_DIVSUB, slot(r, c), slot(a, ca), slot(a, la), slot(b, cb), slot(b, lb)))
                elif kind == MOP_DIV:
                    mops.append((MOP_DIV, slot(r, c), slot(a, ca), slot(a, la), 0, 0))
                else:
                    mops.append((MOP_NEGDIV, slot(r, c), slot(b, cb), slot(b, lb), 0, 0))
        elif op[0] == "reduce":
            s, lt, ls, pairs = red[op[1]]
            t = op[2]
            mops.append((MOP_FACTOR_DIV, slot(t, lt), slot(s, ls), 0, 0, 0))
            mops.append((MOP_ZERO, slot(t, lt), 0, 0, 0, 0))
            pattern[t].add(lt)
            for tc, sc in pairs:
                pattern[t].add(tc)
                mops.append((MOP_SUBMUL, slot(t, tc), slot(s, sc), 0, 0, 0))
        elif op[0] == "rowsub":
            _, t, s, fr, fc = op
            mops.append((MOP_FACTOR_LOAD, slot(fr, fc), 0, 0, 0, 0))
            # Columns outside pattern[s] subtract factor*0 (no change for finite
            # factor); columns in pattern[s] but not yet in pattern[t] become live.
            for c in sorted(pattern[s]):
                pattern[t].add(c)
                mops.append((MOP_SUBMUL, slot(t, c), slot(s, c), 0, 0, 0))
        elif op[0] == "rowscale":
            _, t, fr, fc = op
            mops.append((MOP_FACTOR_INV, slot(fr, fc), 0, 0, 0, 0))
            for c in sorted(pattern[t]):
                mops.append((MOP_SCALE, slot(t, c), 0, 0, 0, 0))
    return slot_of, mops


def run_microops(init, slot_of, mops, f, v, p):
    S = np.zeros(len(slot_of), dtype=np.float64)
    src = (f, v, p)
    for row, col, terms in init:
        acc = None
        for coef, kind, i, j in terms:
            t = np.float64(coef) * src[kind][i, j]
            acc = t if acc is None else acc + t
        S[slot_of[(row, col)]] = acc
    factor = np.float64(0)
    for (o, a, b, c, d, e) in mops:
        if o == MOP_DIVSUB:
            S[a] = S[b] / S[c] - S[d] / S[e]
        elif o == MOP_DIV:
            S[a] = S[b] / S[c]
        elif o == MOP_NEGDIV:
            S[a] = -S[b] / S[c]
        elif o == MOP_FACTOR_DIV:
            factor = S[a] / S[b]
        elif o == MOP_ZERO:
            S[a] = 0.0
        elif o == MOP_SUBMUL:
            S[a] = S[a] - factor * S[b]
        elif o == MOP_FACTOR_LOAD:
            factor = S[a]
        elif o == MOP_FACTOR_INV:
            factor = 1.0 / S[a]
        elif o == MOP_SCALE:
            S[a] = factor * S[a]
    return S


def main():
    init = parse_init()
    spolys = parse_spolys()
    red = parse_reductors()
    code = parse_code()
    assert len(init) == 144 or len(init) > 100, len(init)
    slot_of, mops = build_microops(init, spolys, red, code)
    # The action matrix reads rows 36..41, cols 77..84 (main.cpp:387-390).
    action = []
    for r in range(36, 42):
        for c in range(77, 85):
            action.append(slot_of.get((r, c), -1))
    # --- self-check against the literal dense execution --------------------
    rng = np.random.default_rng(7)
    with np.errstate(all="ignore"):
        for trial in range(5):
            f = rng.standard_normal((3, 3))
            f /= np.linalg.norm(f, axis=0)
            v = rng.standard_normal((3, 3)) * 0.3
            p = rng.standard_normal((3, 3)) * 4
            M = dense_run(init, spolys, red, code, f, v, p)
            S = run_microops(init, slot_of, mops, f, v, p)
            for (r, c), s in slot_of.items():
                a, b = M[r, c], S[s]
                assert (a == b) or (np.isnan(a) and np.isnan(b)), (trial, r, c, a, b)
            # structural zeros really are zero in the dense run
            for r in range(48):
                for c in range(85):
                    if (r, c) not in slot_of:
                        assert M[r, c] == 0.0, (r, c, M[r, c])
    print("self-check OK: %d slots, %d micro-ops, %d init entries" % (len(slot_of), len(mops), len(init)))

    lines = []
    lines.append("// GENERATED by oracle/gen_gp3p_program.py from the opengv GP3P elimination template")
    lines.append("// (dependencies/3rdparty/opengv/src/absolute_pose/modules/gp3p/{init,code,reductors,spolynomials}.cpp).")
    lines.append("// Do not edit. Micro-op table over a dense numbering of structurally non-zero entries.")
    lines.append("#define GP3P_NUM_SLOTS %d" % len(slot_of))
    lines.append("#define GP3P_NUM_INIT %d" % len(init))
    lines.append("#define GP3P_NUM_MOPS %d" % len(mops))
    lines.append("// init: slot, nterms, then 4 x (coef, kind[0=f,1=v,2=p], i, j)")
    lines.append("static const short GP3P_INIT[GP3P_NUM_INIT][18] = {")
    for row, col, terms in init:
        flat = [slot_of[(row, col)], len(terms)]
        for t in terms:
            flat += list(t)
        flat += [0] * (18 - len(flat))
        lines.append("  {" + ",".join(str(x) for x in flat) + "},")
    lines.append("};")
    lines.append("// micro-ops: opcode, a, b, c, d, e (see MOP_* in gen_gp3p_program.py)")
    lines.append("static const short GP3P_MOPS[GP3P_NUM_MOPS][6] = {")
    for m in mops:
        lines.append("  {" + ",".join(str(x) for x in m) + "},")
    lines.append("};")
    lines.append("// slots of groebnerMatrix.block<6,8>(36,77), row-major (-1 = structural zero)")
    lines.append("static const short GP3P_ACTION[48] = {" + ",".join(str(x) for x in action) + "};")
    text = "\n".join(lines) + "\n"
    for o in OUTS:
        os.makedirs(os.path.dirname(o), exist_ok=True)
        with open(o, "w") as fh:
            fh.write(text)
        print("wrote", o, len(text), "bytes")


if __name__ == "__main__":
    sys.exit(main())
