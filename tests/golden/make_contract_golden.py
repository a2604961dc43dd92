"""Extracts the rasterizer CONTRACT from the reference's own source with an `ast` walk and writes
tests/golden/render_contract.json (committed: the GPU box has no /root/reference).

What is read, and from where (all under /root/reference/gs-simp):
  gaussian_renderer/__init__.py   the names imported from `diff_gaussian_rasterization` (:14), the keyword names of the
                                  GaussianRasterizationSettings(...) construction (:36-49), of GaussianRasterizer(...)
                                  (:51) and of the rasterizer(...) call (:85-93), how many values that call is unpacked
                                  into (:85), and the keys of the dict render() returns (:97-101)
  every other *.py                which keys callers read from render()'s result (`render(...)["k"]`, or `pkg["k"]`
                                  for a `pkg = render(...)`), and every comparison of a depth map against a float
                                  constant (the 15.0 sentinel: gen_seq.py:50, vis_render.py:45)
tests/test_abi.py and tests/test_api_gpu.py take the names from the JSON; when /root/reference is present
(this container) tests/test_abi.py re-runs the extraction and requires the committed file to be current.

usage: python tests/golden/make_contract_golden.py [--check]
"""
import ast
import json
import os
import sys

REF = "/root/reference/gs-simp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "render_contract.json")


def _kw(call):
    return [k.arg for k in call.keywords if k.arg is not None]


def _name(node):
    if isinstance(node, ast.Name):
        return node.id
    if isinstance(node, ast.Attribute):
        return node.attr
    return None


def extract_render(path):
    tree = ast.parse(open(path).read(), path)
    out = {"file": os.path.relpath(path, "/root/reference")}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module == "diff_gaussian_rasterization":
            out["imports"] = sorted(a.name for a in node.names)
            out["imports_line"] = node.lineno
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "render")
    out["render_args"] = [a.arg for a in fn.args.args]
    rasterizer_var = None
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call):
            callee = _name(node.value.func)
            if callee == "GaussianRasterizationSettings":
                out["settings_keywords"] = _kw(node.value)
                out["settings_lines"] = [node.value.lineno, node.value.end_lineno]
            elif callee == "GaussianRasterizer":
                out["rasterizer_ctor_keywords"] = _kw(node.value)
                rasterizer_var = node.targets[0].id
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) and _name(node.value.func) == rasterizer_var:
            out["call_keywords"] = _kw(node.value)
            out["call_lines"] = [node.value.lineno, node.value.end_lineno]
            tgt = node.targets[0]
            out["call_returns"] = [e.id for e in tgt.elts] if isinstance(tgt, ast.Tuple) else [tgt.id]
        if isinstance(node, ast.Return) and isinstance(node.value, ast.Dict):
            out["result_keys"] = [k.value for k in node.value.keys]
            out["result_lines"] = [node.value.lineno, node.value.end_lineno]
            vis = [v for k, v in zip(node.value.keys, node.value.values) if k.value == "visibility_filter"]
            if vis:
                out["visibility_filter_expr"] = ast.unparse(vis[0])
    return out


def extract_consumers(root):
    keys, sentinels = {}, []
    for dirpath, _dirs, files in os.walk(root):
        if "SIBR" in dirpath or "submodules" in dirpath:
            continue
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            try:
                tree = ast.parse(open(path).read(), path)
            except SyntaxError:
                continue
            rel = os.path.relpath(path, "/root/reference")
            pkgs = set()
            for node in ast.walk(tree):
                if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) and _name(node.value.func) == "render":
                    for t in node.targets:
                        if isinstance(t, ast.Name):
                            pkgs.add(t.id)
            for node in ast.walk(tree):
                if isinstance(node, ast.Subscript) and isinstance(node.slice, ast.Constant) and isinstance(node.slice.value, str):
                    base = node.value
                    hit = (isinstance(base, ast.Call) and _name(base.func) == "render") or \
                          (isinstance(base, ast.Name) and base.id in pkgs)
                    if hit:
                        keys.setdefault(node.slice.value, []).append(f"{rel}:{node.lineno}")
                if isinstance(node, ast.Compare) and len(node.comparators) == 1:
                    c = node.comparators[0]
                    left = ast.unparse(node.left)
                    if isinstance(c, ast.Constant) and isinstance(c.value, (int, float)) and "depth" in left.lower() \
                            and float(c.value) >= 10.0:
                        sentinels.append({"where": f"{rel}:{node.lineno}", "expr": ast.unparse(node), "value": float(c.value)})
    return {"result_keys_read": {k: sorted(v) for k, v in sorted(keys.items())}, "depth_sentinel_compares": sentinels}


def build():
    c = extract_render(os.path.join(REF, "gaussian_renderer", "__init__.py"))
    c.update(extract_consumers(REF))
    return c


if __name__ == "__main__":
    c = build()
    text = json.dumps(c, indent=1, sort_keys=True) + "\n"
    if "--check" in sys.argv:
        assert open(OUT).read() == text, "tests/golden/render_contract.json is stale: re-run make_contract_golden.py"
        print("render_contract.json is current")
    else:
        open(OUT, "w").write(text)
        print(text)
