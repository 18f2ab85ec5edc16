"""Builds tests/golden/prng_golden.json: the 100-bit known answer of the TS 38.211 5.2.1 Gold sequence typed into the
reference's unit test (/root/reference/test/unit/nr/test_nr_utils.py:280-296, n_rnti = 20001, n_id = 41)."""
import ast
import json
import os

SRC = "/root/reference/test/unit/nr/test_nr_utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prng_golden.json")

tree = ast.parse(open(SRC).read())
vals = {}
for fn in ast.walk(tree):
    if isinstance(fn, ast.FunctionDef) and fn.name == "test_gen_rand_seq":
        for st in fn.body:
            if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name):
                name = st.targets[0].id
                if name in ("n_rnti", "n_id", "l"):
                    vals[name] = ast.literal_eval(st.value)
                if name == "s_ref":
                    vals["s_ref"] = [int(v) for v in ast.literal_eval(st.value.args[0])]
with open(OUT, "w") as f:
    json.dump(vals, f)
print(vals["n_rnti"], vals["n_id"], vals["l"], len(vals["s_ref"]))
