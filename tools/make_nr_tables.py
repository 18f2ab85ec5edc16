"""Extracts 3GPP table DATA needed by sionna_b200.phy.nr from the reference checkout into
sionna_b200/phy/nr/codes/nr_tables.npz (run in the build container, /root/reference present):

  * MCS tables of TS 38.214 (5.1.3.1-1..4 and 6.1.4.1-1/2): modulation orders and target rates x1024
    (literal lists inside decode_mcs_index, /root/reference/src/sionna/phy/nr/utils.py:175-240)
  * PUSCH codebooks W of TS 38.211 Tables 6.3.1.5-1..7, one array per (layers, antenna ports)
    (PUSCHConfig.precoding_matrix, /root/reference/src/sionna/phy/nr/pusch_config.py:597-807)

Only numbers are stored. The reference cannot be imported (TensorFlow is absent), so the two code fragments are
located with `ast` and evaluated in isolation.
"""
import ast
import os
import types
import numpy as np

REF = "/root/reference/src/sionna/phy/nr"
OUT = os.path.join(os.path.dirname(__file__), "..", "sionna_b200", "phy", "nr", "codes", "nr_tables.npz")


def literal_lists(path, func, names):
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            for st in ast.walk(node):
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and st.targets[0].id in names:
                    call = st.value                          # tf.convert_to_tensor([...])
                    arg = call.args[0] if isinstance(call, ast.Call) else call
                    out[st.targets[0].id] = eval(compile(ast.Expression(arg), "<mcs>", "eval"), {})
    return out


def precoding_tables(path):
    tree = ast.parse(open(path).read())
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "precoding_matrix":
            fn = node
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
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
