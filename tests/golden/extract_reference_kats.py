#!/usr/bin/env python3
"""Mechanical provenance of tests/golden/reference_kats.json.

The reference cannot be compiled here (no rustc), so its known-answer vectors were transcribed by hand in
round 1.  This script re-derives them from the Rust sources under /root/reference instead of trusting the
transcription:

  * the rstest `#[case(...)]` tables of src/algebra/field/prime/arithmetic.rs (add, sub, mul, field_pow,
    multiplicative_inverse, halve) are PARSED and must contain every entry of the JSON's field section
    (`--write` regenerates that section from the parse, adding nothing by hand);
  * every other integer vector of the JSON (polynomial, GF(101²), curve, kzg, Reed–Solomon sections) is located
    in the cited source file as a contiguous run of its numeric literals, after stripping the type parameters
    (`PlutoBaseField`, const-generic sizes, `usize` suffixes) — a vector that cannot be found is an error;
  * sections the JSON itself marks as derived (config1_extra: computed with the reference's schoolbook algorithm,
    reed_solomon_decode: round trips) are listed as derived, not searched.

Run:  python tests/golden/extract_reference_kats.py            (check; exit code 1 on any unlocated vector)
      python tests/golden/extract_reference_kats.py --write    (also rewrite the parsed sections in place)
tests/test_oracle_golden.py runs the check when /root/reference is present (it is absent on the GPU box)."""
from __future__ import annotations

import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RONK_REFERENCE", "/root/reference")
JSON_PATH = os.path.join(HERE, "reference_kats.json")
FIELD_OF = {"PlutoScalarField": 17, "PlutoBaseField": 101}


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def rstest_cases(src: str, fn_name: str):
    """The `#[case(...)]` lines directly above `fn <fn_name>`; each case → list of (modulus | None, int)."""
    m = re.search(r"((?:\s*(?://[^\n]*|#\[[^\n]*\])\n)+)\s*fn " + re.escape(fn_name) + r"\b", src)
    assert m, f"fn {fn_name} not found"
    cases = []
    for line in m.group(1).splitlines():
        line = line.strip()
        if not line.startswith("#[case("):
            continue
        toks = re.findall(r"(Pluto(?:Scalar|Base)Field)::new\((\d+)\)|(?<![\w:])(\d+)(?![\w:])", line)
        vals = []
        for fld, v, bare in toks:
            vals.append((FIELD_OF[fld], int(v)) if fld else (None, int(bare)))
        cases.append(vals)
    return cases


def field_section_from_source():
    src = read("src/algebra/field/prime/arithmetic.rs")
    out = {}
    for key, fn in (("add", "add"), ("sub", "sub"), ("mul", "mul")):
        out[key] = [[c[0][0], c[0][1], c[1][1], c[2][1]] for c in rstest_cases(src, fn)]
    out["pow"] = [[c[0][0], c[0][1], c[1][1], c[2][1]] for c in rstest_cases(src, "field_pow")]
    inv = rstest_cases(src, "multiplicative_inverse")
    out["inverse"] = [[c[0][0], c[0][1], c[1][1]] for c in inv if c[0][1] != 0]       # the 0 cases are #[should_panic]
    out["inverse_of_zero_panics"] = sorted({c[0][0] for c in inv if c[0][1] == 0})
    out["halve"] = [[c[0][0], c[0][1], c[1][1]] for c in rstest_cases(src, "halve")]
    return out


def number_stream(src: str):
    """Numeric literals of a Rust source in order, without const-generic sizes / type parameters / suffixes."""
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"::<\{[^}]*\}>", "", src)                    # PrimeField::<{ PlutoPrime::Base as usize }>
    src = re.sub(r"::<[^>]*>", "", src)                       # Polynomial::<Monomial, PlutoBaseField, 4>
    src = re.sub(r"\[([^\[\];]+);\s*(\d+)\]", lambda m: "[" + ", ".join([m.group(1)] * int(m.group(2))) + "]", src)  # [x; n]
    src = re.sub(r"(?<=\w)\[\s*\d+\s*\]", "", src)            # index expressions: arr[0], data[1]
    src = re.sub(r"<[A-Za-z_][\w, ]*\d+\s*>", "", src)        # Polynomial<Monomial, PlutoBaseField, 4>
    src = src.replace("::ZERO", "::new(0)").replace("::ONE", "::new(1)")
    return [int(x) for x in re.findall(r"(?<![\w.])(\d+)(?:usize|u32|u64|i32)?(?![\w.])", src)]


def contains_run(stream, vec):
    n = len(vec)
    if n == 0:
        return True
    first = vec[0]
    for i, v in enumerate(stream):
        if v == first and stream[i:i + n] == vec:
            return True
    return False


def point_literals(p):
    """[x0,x1,y0,y1] → the literals the reference writes: base-field points as (x0, y0), extension points with
    all four coordinates (in either order the sources use)."""
    x0, x1, y0, y1 = p
    if x1 == 0 and y1 == 0:
        return [[x0, y0]]
    return [[x0, x1, y0, y1], [x0, y0, y1], [x0, y1], [x0, 0, 0, y1]]


def main():
    write = "--write" in sys.argv
    with open(JSON_PATH) as f:
        kats = json.load(f)
    problems, located, derived = [], 0, []

    # ---- field section: parsed tables -------------------------------------------------------------------------
    parsed = field_section_from_source()
    for key, rows in parsed.items():
        have = kats["field"].get(key)
        if have is None:
            continue
        for row in have:
            if row not in rows:
                problems.append(f"field.{key}: {row} is not a #[case] of the reference")
            else:
                located += 1
    if write:
        for key, rows in parsed.items():
            kats["field"][key] = rows

    # ---- every other vector: located in the cited file ----------------------------------------------------------
    streams = {}

    def stream(rel):
        if rel not in streams:
            streams[rel] = number_stream(read(rel))
        return streams[rel]

    def locate(label, vec, files):
        nonlocal located
        vec = [int(v) for v in vec]
        if any(contains_run(stream(f), vec) for f in files):
            located += 1
        else:
            problems.append(f"{label}: {vec} not found in {files}")

    poly_files = ["src/polynomial/tests.rs", "src/polynomial/arithmetic.rs"]
    for key, v in kats["polynomial"].items():
        if isinstance(v, list) and v and all(isinstance(x, int) for x in v):
            locate(f"polynomial.{key}", v, poly_files)
    gf_file = ["src/algebra/field/extension/gf_101_2.rs"]
    for op, rows in kats["gf101_2"].items():
        if op == "src":
            continue
        for row in rows:
            for pair in row:
                locate(f"gf101_2.{op}", pair, gf_file)
    curve_files = ["src/curve/pluto_curve.rs", "src/kzg/tests.rs", "src/kzg/setup.rs"]
    pts = [kats["curve"]["G1"], kats["curve"]["G2"], kats["curve"]["two_G2"], kats["curve"]["off_curve"]]
    pts += list(kats["curve"]["multiples_of_G1"].values()) + kats["kzg"]["g1srs"] + kats["kzg"]["g2srs"]
    for p in pts:
        forms = point_literals(p)
        if any(contains_run(stream(f), form) for f in curve_files for form in forms):
            located += 1
        else:
            problems.append(f"curve/kzg point {p}: none of {forms} found in {curve_files}")
    for c in kats["kzg"]["commit"]:
        locate("kzg.commit.coeffs", c["coeffs"], ["src/kzg/tests.rs"])
    rs = kats["reed_solomon"]
    for key in ("msg", "x", "y"):
        locate(f"reed_solomon.{key}", rs[key], ["src/codes/reed_solomon.rs"])
    for key in ("config1_extra", "reed_solomon_decode"):
        if key in kats:
            derived.append(key)

    if write:
        with open(JSON_PATH, "w") as f:
            json.dump(kats, f, indent=1)
            f.write("\n")
    print(json.dumps({"located": located, "derived_sections": derived, "problems": problems}, indent=1))
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
