"""Poor man's pyflakes (no linter in the image): import every module of the package (+ bench.py, tests) and report
LOAD_GLOBAL names that resolve neither in the module namespace nor in builtins.  Catches NameErrors before a GPU run."""
import builtins
import dis
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def code_objects(co):
    yield co
    for c in co.co_consts:
        if isinstance(c, types.CodeType):
            yield from code_objects(c)


def check(modname=None, path=None):
    if path:
        src = open(path).read()
        co = compile(src, path, 'exec')
        ns = {'__name__': '__lint__', '__file__': path}
        # collect module-level names without executing: assigned names + imports + defs
        import ast
        names = set()
        for node in ast.walk(ast.parse(src)):
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                names.add(node.name)
            elif isinstance(node, ast.Import):
                names.update((a.asname or a.name).split('.')[0] for a in node.names)
            elif isinstance(node, ast.ImportFrom):
                names.update(a.asname or a.name for a in node.names)
            elif isinstance(node, ast.Name) and isinstance(node.ctx, ast.Store):
                names.add(node.id)
            elif isinstance(node, ast.Global):
                names.update(node.names)
            elif isinstance(node, ast.arg):
                names.add(node.arg)
        have = lambda n: n in names or hasattr(builtins, n) or n in ('__file__', '__name__')
        label = path
    else:
        mod = importlib.import_module(modname)
        co = compile(open(mod.__file__).read(), mod.__file__, 'exec')
        have = lambda n: hasattr(mod, n) or hasattr(builtins, n)
        label = modname
    bad = []
    for c in code_objects(co):
        for ins in dis.get_instructions(c):
            if ins.opname in ('LOAD_GLOBAL', 'LOAD_NAME') and not have(ins.argval) and not (
                    ins.opname == 'LOAD_NAME' and ins.argval in c.co_names and c.co_name != '<module>'):
                bad.append((c.co_name, ins.argval, ins.positions.lineno if ins.positions else None))
    for b in sorted(set(bad)):
        print(f'{label}: undefined name {b[1]!r} in {b[0]} (line {b[2]})')
    return len(bad)


if __name__ == '__main__':
    n = 0
    pkg = 'clip_decontamination_b200'
    for root, _, files in os.walk(os.path.join(ROOT, pkg)):
        for f in files:
            if f.endswith('.py'):
                rel = os.path.relpath(os.path.join(root, f), ROOT)[:-3].replace(os.sep, '.')
                if rel.endswith('.__init__'):
                    rel = rel[:-9]
                n += check(modname=rel)
    for p in ['bench.py', '__graft_entry__.py', 'oracle/ref_on_gpu.py', 'oracle/gen_golden.py', 'oracle/ref_harness.py'] + \
            [os.path.join('tests', f) for f in sorted(os.listdir(os.path.join(ROOT, 'tests'))) if f.endswith('.py')]:
        n += check(path=os.path.join(ROOT, p))
    print('undefined names:', n)
    sys.exit(1 if n else 0)
