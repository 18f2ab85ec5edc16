"""Builds tests/golden/mcs_golden.json from the MCS known-answer lists typed into the reference's unit tests
(/root/reference/test/unit/nr/test_nr_utils.py: test_mcs_pdsch :17-102, test_mcs_pusch :104-268, values of TS 38.214
Tables 5.1.3.1-1..4 and 6.1.4.1-1/2). The test source is parsed with `ast`; every (qs, rs) pair is stored with the keyword
arguments of the decode_mcs_index call that follows it. Run where /root/reference exists."""
import ast
import json
import os

SRC = "/root/reference/test/unit/nr/test_nr_utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mcs_golden.json")


def const(node, env):
    if isinstance(node, ast.Constant):
        return node.value
    if isinstance(node, ast.Name):
        return env.get(node.id, "both" if node.id == "bpsk" else None)
    raise ValueError(ast.dump(node))


def main():
    tree = ast.parse(open(SRC).read())
    cases = []
    for fn in ast.walk(tree):
        if not (isinstance(fn, ast.FunctionDef) and fn.name in ("test_mcs_pdsch", "test_mcs_pusch")):
            continue
        env, pending = {}, None
        for st in fn.body:
            if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name):
                name = st.targets[0].id
                if name in ("qs", "rs"):
                    env[name] = ast.literal_eval(st.value)
                    if name == "rs":
                        pending = {"qs": env["qs"], "rs": env["rs"]}
                elif name == "pi2bpsk":
                    env[name] = ast.literal_eval(st.value)
            elif pending is not None and isinstance(st, (ast.For, ast.With)):
                for call in ast.walk(st):
                    if isinstance(call, ast.Call) and getattr(call.func, "id", "") == "decode_mcs_index":
                        kw = {k.arg: const(k.value, env) for k in call.keywords if k.arg != "mcs_index"}
                        if isinstance(st, ast.For):
                            cases.append({**pending, "kwargs": kw, "test": fn.name})
                            pending = None
                        break
    with open(OUT, "w") as f:
        json.dump(cases, f)
    for c in cases:
        print(c["test"], c["kwargs"], len(c["qs"]))


if __name__ == "__main__":
    main()
