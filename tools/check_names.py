# poor man's pyflakes: undefined global names in test functions
import ast, sys, builtins
for path in sys.argv[1:]:
    tree = ast.parse(open(path).read())
    defined = set(dir(builtins))
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names: defined.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, (ast.FunctionDef, ast.ClassDef)): defined.add(node.name)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, ast.Store): defined.add(node.id)
        elif isinstance(node, ast.arg): defined.add(node.arg)
        elif isinstance(node, ast.ExceptHandler) and node.name: defined.add(node.name)
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load) and node.id not in defined:
            print(path, node.lineno, node.id)
