"""Builds tests/golden/layer_mapper_golden.json: the predefined input / output sequences of the reference's LayerMapper
test for 1..8 layers (/root/reference/test/unit/nr/test_layer_mapper.py:14-203, single and dual codeword mode)."""
import ast
import json
import os

SRC = "/root/reference/test/unit/nr/test_layer_mapper.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layer_mapper_golden.json")

tree = ast.parse(open(SRC).read())
env, cases = {}, []
for fn in ast.walk(tree):
    if isinstance(fn, ast.FunctionDef) and fn.name == "test_ref":
        for st in fn.body:
            if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and isinstance(st.value, ast.Call) \
                    and getattr(st.value.func, "attr", "") == "array":
                env[st.targets[0].id] = ast.literal_eval(st.value.args[0])
for layers in range(1, 9):
    out = env[f"o{layers}"]
    ins = [env["u"]] if layers <= 4 else None
    cases.append({"num_layers": layers, "out": out, "inputs": ins})
# dual-codeword inputs are re-assigned (u1, u2) before every case: walk again in order
order = []
for fn in ast.walk(tree):
    if isinstance(fn, ast.FunctionDef) and fn.name == "test_ref":
        cur = {}
        for st in fn.body:
            if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and isinstance(st.value, ast.Call) \
                    and getattr(st.value.func, "attr", "") == "array":
                name = st.targets[0].id
                cur[name] = ast.literal_eval(st.value.args[0])
                if name in ("o5", "o6", "o7", "o8"):
                    cases[int(name[1]) - 1]["inputs"] = [cur["u1"], cur["u2"]]
with open(OUT, "w") as f:
    json.dump(cases, f)
for c in cases:
    print(c["num_layers"], [len(i[0]) for i in c["inputs"]], len(c["out"][0]), len(c["out"][0][0]))
