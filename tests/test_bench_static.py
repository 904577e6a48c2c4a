"""bench.py runs on the GPU box only; a name that does not resolve would surface there, at the end of a round.  This walks its
syntax tree on the CPU: every global name a function reads must be defined at module level, imported, or a builtin."""
import ast
import builtins
import os
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _unresolved(path):
    src = open(path).read()
    tree = ast.parse(src)
    module_names = set(dir(builtins))
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            module_names.add(node.name) if node in tree.body else None
        elif isinstance(node, (ast.Import, ast.ImportFrom)) and node in tree.body:
            for a in node.names:
                module_names.add((a.asname or a.name).split(".")[0])
    for node in tree.body:
        if isinstance(node, ast.Assign):
            for t in node.targets:
                for n in ast.walk(t):
                    if isinstance(n, ast.Name):
                        module_names.add(n.id)
    bad = []

    def visit(tab):
        for sym in tab.get_symbols():
            if tab.get_type() == "function" and sym.is_global() and sym.is_referenced() and not sym.is_assigned() and sym.get_name() not in module_names:
                bad.append((tab.get_name(), sym.get_name()))
        for child in tab.get_children():
            visit(child)
    visit(symtable.symtable(src, path, "exec"))
    return bad


def test_bench_names_resolve():
    assert _unresolved(os.path.join(ROOT, "bench.py")) == []


def test_entry_names_resolve():
    assert _unresolved(os.path.join(ROOT, "__graft_entry__.py")) == []
