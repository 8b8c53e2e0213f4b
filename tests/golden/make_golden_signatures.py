#!/usr/bin/env python
"""Extracts the Python-level signatures of the reference's hot-path wrappers (names, order, defaults) by parsing
their sources with `ast` -- no import, the reference package needs `future`/`wurlitzer` -- into
tests/golden/reference_signatures.json.  Run in the build container only."""
import ast
import json
import os

REF = "/root/reference/Corrfunc"
FUNCS = {"DD": "theory/DD.py", "DDrppi": "theory/DDrppi.py", "DDsmu": "theory/DDsmu.py", "wp": "theory/wp.py",
         "xi": "theory/xi.py", "vpf": "theory/vpf.py", "DDtheta_mocks": "mocks/DDtheta_mocks.py", "DDrppi_mocks": "mocks/DDrppi_mocks.py",
         "DDsmu_mocks": "mocks/DDsmu_mocks.py", "vpf_mocks": "mocks/vpf_mocks.py",
         "convert_3d_counts_to_cf": "utils.py", "convert_rp_pi_counts_to_wp": "utils.py",
         "return_file_with_rbins": "utils.py", "fix_cz": "utils.py", "fix_ra_dec": "utils.py",
         "translate_isa_string_to_enum": "utils.py", "compute_nbins": "utils.py", "gridlink_sphere": "utils.py",
         "convert_to_native_endian": "utils.py", "is_native_endian": "utils.py", "process_weights": "utils.py"}
out = {}
for name, rel in FUNCS.items():
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    args = [a.arg for a in fn.args.args]
    defaults = [ast.literal_eval(d) for d in fn.args.defaults]
    nreq = len(args) - len(defaults)
    out[name] = {"source": "Corrfunc/%s:%d" % (rel, fn.lineno), "args": args,
                 "defaults": dict(zip(args[nreq:], defaults))}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_signatures.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path)
for k, v in out.items():
    print(k, v["args"])
