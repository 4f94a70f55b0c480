"""Extracts the argparse surface of the UNMODIFIED reference CLIs (/root/reference/main_train.py, main_eval.py) by
parsing their source (no import: they need thop / CUDA) and writes tests/golden/cli_flags.json.
Run in the build container:  python tests/golden/make_cli_golden.py"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def flags_of(path):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == 'add_argument':
            names = [a.value for a in node.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
            kw = {}
            for k in node.keywords:
                if k.arg in ('default', 'nargs', 'action', 'dest'):
                    try:
                        kw[k.arg] = ast.literal_eval(k.value)
                    except Exception:
                        kw[k.arg] = ast.unparse(k.value)
                elif k.arg == 'type':
                    kw['type'] = ast.unparse(k.value)
            out.append({'flags': names, **kw})
    return out


if __name__ == '__main__':
    data = {name: flags_of(os.path.join('/root/reference', name)) for name in ('main_train.py', 'main_eval.py')}
    with open(os.path.join(HERE, 'cli_flags.json'), 'w') as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in data.items()})
