#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ from the reference's own test data.

Run ONCE in the authoring container (where /root/reference exists); the outputs are committed.
Nothing at test time reads /root/reference.

Sources (cpmech/russell @ 44fc3f9):
  * russell_sparse/src/samples.rs            -> samples.json  (COO triplets + CSC + CSR + det per sample/variant)
                                             -> complex_samples.json (the Complex* twins; values as [re, im] pairs)
  * russell_sparse/src/bin/solve_matrix_market.rs:307-372 -> bfwb62_x.json (62 golden solution values)
  * russell_sparse/data/matrix_market/*.mtx  -> copied verbatim as parser fixtures (data files, not code)
"""
import itertools, json, os, re, shutil, sys

REF = "/root/reference/russell_sparse"
OUT = os.path.dirname(os.path.abspath(__file__))


def eval_num(expr):
    expr = expr.strip()
    if not re.fullmatch(r"[-+*/ ().0-9eE]+", expr):
        raise ValueError("unexpected numeric expression: %r" % expr)
    return float(eval(expr))


def eval_val(expr):
    """real literal -> float; cpx!(a, b) -> [a, b]"""
    expr = expr.strip()
    m = re.fullmatch(r"cpx!\((.+),(.+)\)", expr)
    if m:
        return [eval_num(m.group(1)), eval_num(m.group(2))]
    return eval_num(expr)


def split_items(buf):
    """splits a vec![...] body at top-level commas (cpx!(a, b) holds a comma of its own)"""
    items, depth, cur = [], 0, ""
    for ch in buf:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            items.append(cur)
            cur = ""
        else:
            cur += ch
    items.append(cur)
    return [t.strip() for t in items if t.strip()]


def strip_comment(line):
    k = line.find("//")
    return line if k < 0 else line[:k]


def parse_samples(want_complex=False):
    src = open(os.path.join(REF, "src/samples.rs")).read().split("\n")
    # locate real-valued sample functions
    starts = [i for i, l in enumerate(src) if re.match(r"\s*pub fn \w+\(", l)]
    out = {}
    for si, s in enumerate(starts):
        end = starts[si + 1] if si + 1 < len(starts) else len(src)
        # function header may span several lines up to '{'
        hdr = ""
        k = s
        while "{" not in src[k]:
            hdr += src[k]
            k += 1
        hdr += src[k]
        name = re.search(r"pub fn (\w+)\(", hdr).group(1)
        if ("Complex" in hdr) != want_complex:
            continue
        params = re.findall(r"(\w+): bool", hdr)
        body = src[k + 1:end]
        for combo in itertools.product([True, False], repeat=len(params)):
            env = dict(zip(params, combo))
            rec = run_body(body, env)
            if rec is None:
                continue
            key = name if not params else name + "(" + ",".join(str(env[p]).lower() for p in params) + ")"
            out[key] = rec
    return out


def run_body(body, env):
    """Tiny interpreter for the regular shape of the sample functions (if/else on bool params)."""
    rec = {"coo_i": [], "coo_j": [], "coo_v": []}
    active = [True]  # stack of branch activity
    kinds = []       # 'if' or 'other' per open brace
    depth0_done = False
    vec_name, vec_buf = None, None
    which = None  # 'csc' or 'csr' values disambiguation
    i = 0
    while i < len(body):
        line = strip_comment(body[i]).strip()
        i += 1
        if not line:
            continue
        if vec_name is not None:
            vec_buf += " " + line
            if "]" in line:
                finish_vec(rec, vec_name, vec_buf)
                vec_name = None
            continue
        m = re.match(r"if (!?)(\w+) \{$", line)
        if m:
            val = env[m.group(2)]
            if m.group(1):
                val = not val
            active.append(active[-1] and val)
            kinds.append(("if", val))
            continue
        if line == "} else {":
            k, val = kinds[-1]
            active.pop()
            active.append(active[-1] and (not val))
            continue
        if line == "}":
            if not kinds:
                break  # end of function
            kinds.pop()
            active.pop()
            continue
        if not active[-1]:
            continue
        m = re.match(r"let sym = Sym::(\w+);", line)
        if m:
            rec["sym"] = m.group(1)
            continue
        m = re.match(r"let (nrow|ncol|max_nnz) = (\d+);", line)
        if m:
            rec[m.group(1)] = int(m.group(2))
            continue
        m = re.match(r"let \(nrow, ncol, (?:nnz|max_nnz)\) = \((\d+), (\d+), (\d+)\);", line)
        if m:
            rec["nrow"], rec["ncol"], rec["max_nnz"] = int(m.group(1)), int(m.group(2)), int(m.group(3))
            continue
        m = re.match(r"coo\.put\((\d+), (\d+), (.+)\)\.unwrap\(\);", line)
        if m:
            rec["coo_i"].append(int(m.group(1)))
            rec["coo_j"].append(int(m.group(2)))
            rec["coo_v"].append(eval_val(m.group(3)))
            continue
        m = re.match(r"let (values|row_indices|col_indices|col_pointers|row_pointers) = vec!\[(.*)$", line)
        if m:
            vec_name, vec_buf = m.group(1), m.group(2)
            if "]" in vec_buf:
                finish_vec(rec, vec_name, vec_buf)
                vec_name = None
            continue
        m = re.match(r"\(coo, csc, csr, (.+)\)$", line)
        if m:
            try:
                rec["det"] = eval_val(m.group(1))
            except ValueError:
                rec["det"] = None  # rectangular samples carry a placeholder instead of a determinant
            continue
    if "sym" not in rec or "det" not in rec:
        return None
    rec.setdefault("max_nnz", len(rec["coo_v"]))  # some samples pass a literal to CooMatrix::new
    return rec


def finish_vec(rec, name, buf):
    buf = buf[:buf.index("]")]
    items = split_items(buf)
    if name == "values":
        key = "csc_values" if "csc_values" not in rec else "csr_values"
        rec[key] = [eval_val(t) for t in items]
    else:
        rec[name] = [int(t) for t in items]


def parse_bfwb62_x():
    src = open(os.path.join(REF, "src/bin/solve_matrix_market.rs")).read()
    k = src.index("fn get_bfwb62_correct_x")
    blk = src[k:src.index("])", k)]
    vals = [float(t) for t in re.findall(r"(-?\d\.\d+e[+-]\d+)", blk)]
    assert len(vals) == 62, len(vals)
    return vals


def main():
    samples = parse_samples()
    with open(os.path.join(OUT, "samples.json"), "w") as f:
        json.dump(samples, f, indent=1, sort_keys=True)
    csamples = parse_samples(want_complex=True)
    with open(os.path.join(OUT, "complex_samples.json"), "w") as f:
        json.dump(csamples, f, indent=1, sort_keys=True)
    print("complex samples:", len(csamples), sorted(csamples))
    with open(os.path.join(OUT, "bfwb62_x.json"), "w") as f:
        json.dump(parse_bfwb62_x(), f, indent=1)
    mm = os.path.join(OUT, "matrix_market")
    os.makedirs(mm, exist_ok=True)
    for fn in sorted(os.listdir(os.path.join(REF, "data/matrix_market"))):
        if fn.endswith(".mtx"):
            shutil.copy(os.path.join(REF, "data/matrix_market", fn), os.path.join(mm, fn))
    print("samples:", len(samples), sorted(samples)[:5], "...")


if __name__ == "__main__":
    main()
