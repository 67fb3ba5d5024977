"""Extract the 3SPN.2 / 3SPN.2C parameter tables (published values: Hinckley 2013, Freeman 2014, Lu 2021) from the
reference headers into opencafemol_b200/tables_3spn2.json, in the units the reference hands to OpenMM (nm, rad,
kJ/mol).  std::map semantics are reproduced: for a duplicated key the FIRST initialiser wins.

usage: python tools/extract_3spn2_tables.py [/root/reference]      (needs the reference tree; the JSON is committed)"""
import json
import math
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
FF = os.path.join(REF, "src", "forcefield")
CONST = {"OpenMM::NmPerAngstrom": "0.1", "OpenMM::RadiansPerDegree": repr(3.1415926535897932385 / 180.0)}


def evaluate(expr):
    for k, v in CONST.items():
        expr = expr.replace(k, v)
    expr = re.sub(r"double\(([^()]*)\)", r"(\1)", expr)
    if not re.fullmatch(r"[0-9eE+\-*/. ()]+", expr):
        raise ValueError(expr)
    return float(eval(expr))  # arithmetic on literals only (checked above)


def struct_bodies(text):
    for m in re.finditer(r"struct\s+(\w+)[^;{]*\{", text):
        depth, i = 1, m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        yield m.group(1), text[m.end():i - 1]


def maps_in(body):
    out = {}
    # "name = {" (static member) or "name() const override { return std::map<...>{"
    for m in re.finditer(r"(\w+)\s*(?:\(\s*\)\s*const\s*override\s*\{\s*return\s*std::map<std::string,\s*double>\s*|=\s*)\{", body):
        depth, i = 1, m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(body[i], 0)
            i += 1
        block = re.sub(r"//[^\n]*", "", body[m.end():i - 1])
        entries = re.findall(r'\{\s*"([^"]+)"\s*,\s*([^{}]+?)\s*\}', block)
        if not entries:
            continue
        table = {}
        for k, e in entries:
            if k not in table:  # std::map initialiser list: first occurrence wins
                table[k] = evaluate(e)
        out[m.group(1)] = table
    return out


def scalars_in(body):
    out = {}
    for m in re.finditer(r"double\s+(\w+)\s*\(\s*\)\s*const\s*override\s*\{\s*return\s*([^;]+);", body):
        out[m.group(1)] = evaluate(m.group(2).strip())
    for m in re.finditer(r"static\s+const\s+double\s+(\w+)\s*=\s*([^;]+);", body):
        out[m.group(1)] = evaluate(re.sub(r"//.*", "", m.group(2)).strip())
    return out


def main():
    result = {}
    for fname in ("ThreeSPN2DefaultParameters.hpp", "ThreeSPN2BaseStackingForceFieldGenerator.hpp",
                  "ThreeSPN2ExcludedVolumeForceFieldGenerator.hpp"):
        text = open(os.path.join(FF, fname)).read()
        for name, body in struct_bodies(text):
            if not name.endswith("Parameter"):
                continue
            m, s = maps_in(body), scalars_in(body)
            if m or s:
                result[name] = {"maps": m, "scalars": s}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "opencafemol_b200", "tables_3spn2.json")
    with open(out, "w") as fh:
        json.dump(result, fh, indent=1, sort_keys=True)
    for k, v in result.items():
        print(k, {a: len(b) for a, b in v["maps"].items()}, v["scalars"])


if __name__ == "__main__":
    main()
