"""Extracts 3GPP table DATA needed by sionna_b200.phy.nr from the reference checkout into
sionna_b200/phy/nr/codes/nr_tables.npz (run in the build container, /root/reference present):

  * MCS tables of TS 38.214 (5.1.3.1-1..4 and 6.1.4.1-1/2): modulation orders and target rates x1024
    (literal lists inside decode_mcs_index, /root/reference/src/sionna/phy/nr/utils.py:175-240)
  * PUSCH codebooks W of TS 38.211 Tables 6.3.1.5-1..7, one array per (layers, antenna ports)
    (PUSCHConfig.precoding_matrix, /root/reference/src/sionna/phy/nr/pusch_config.py:597-807)

Only numbers are stored. The reference cannot be imported (TensorFlow is absent), so the two code fragments are
located with `ast`: the MCS lists go through a numbers-and-arithmetic-only evaluator; the codebook function is vetted node by node
(NumPy arithmetic only) and executed with a five-name builtins table.
"""
import ast
import os
import types
import numpy as np

REF = "/root/reference/src/sionna/phy/nr"
OUT = os.path.join(os.path.dirname(__file__), "..", "sionna_b200", "phy", "nr", "codes", "nr_tables.npz")


def _num_eval(node):
    """Evaluates a tree of numbers, lists / tuples and + - * / only (what the MCS tables are written with)."""
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return node.value
    if isinstance(node, (ast.List, ast.Tuple)):
        return [_num_eval(e) for e in node.elts]
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = _num_eval(node.operand)
        return -v if isinstance(node.op, ast.USub) else v
    if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Sub, ast.Mult, ast.Div)):
        a, b = _num_eval(node.left), _num_eval(node.right)
        if isinstance(node.op, ast.Add):
            return a + b                                        # numbers, or list concatenation
        if isinstance(node.op, ast.Mult):
            return a * b                                        # numbers, or list repetition
        if isinstance(a, list) or isinstance(b, list):
            raise ValueError("list operand of - or /")
        return a - b if isinstance(node.op, ast.Sub) else a / b
    raise ValueError("unexpected node in a numeric table: " + ast.dump(node)[:80])


def literal_lists(path, func, names):
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            for st in ast.walk(node):
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and st.targets[0].id in names:
                    call = st.value                          # tf.convert_to_tensor([...])
                    arg = call.args[0] if isinstance(call, ast.Call) else call
                    out[st.targets[0].id] = _num_eval(arg)             # nested number lists with simple arithmetic: nothing is executed
    return out


def precoding_tables(path):
    tree = ast.parse(open(path).read())
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "precoding_matrix":
            fn = node
    fn.decorator_list = []
    # The codebooks are written as NumPy expressions (np.array([...]) / np.sqrt(2), 1j factors), not literals, so this one
    # function body is executed - after checking that it contains nothing but arithmetic on `np` and `self` attributes
    # (no imports, no calls other than np.* / complex / float / int / len / range, no attribute access on anything else)
    # and with only those five builtins.
    for node in ast.walk(fn):
        if isinstance(node, (ast.Import, ast.ImportFrom, ast.Global, ast.Nonlocal, ast.Lambda, ast.With, ast.Try)):
            raise RuntimeError("unexpected construct in the reference's precoding_matrix")
        if isinstance(node, ast.Call):
            f = node.func
            ok = (isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "np") or \
                 (isinstance(f, ast.Name) and f.id in ("complex", "float", "int", "len", "range"))
            if not ok:
                raise RuntimeError("unexpected call in the reference's precoding_matrix")
        if isinstance(node, ast.Attribute) and not (isinstance(node.value, ast.Name) and node.value.id in ("np", "self")):
            raise RuntimeError("unexpected attribute access in the reference's precoding_matrix")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "__builtins__": {"complex": complex, "float": float, "int": int, "len": len, "range": range}}
    exec(compile(mod, "<w>", "exec"), ns)
    tabs = {}
    for layers, ports, count in ((1, 2, 6), (1, 4, 28), (2, 2, 3), (2, 4, 22), (3, 4, 7), (4, 4, 5)):
        ws = []
        for tpmi in range(count):
            cfg = types.SimpleNamespace(precoding="codebook", num_layers=layers, num_antenna_ports=ports, tpmi=tpmi)
            ws.append(np.asarray(ns["precoding_matrix"](cfg), complex))
        tabs[f"w_{layers}_{ports}"] = np.stack(ws)
    return tabs


def main():
    mcs = literal_lists(os.path.join(REF, "utils.py"), "decode_mcs_index", ("mod_orders", "target_rates"))
    out = {"mcs_mod_orders": np.array(mcs["mod_orders"], np.int32), "mcs_target_rates": np.array(mcs["target_rates"], np.float64)}
    out.update(precoding_tables(os.path.join(REF, "pusch_config.py")))
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
