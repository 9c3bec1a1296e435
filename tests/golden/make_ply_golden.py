"""tests/golden/ply_golden.pt: the attribute list of the REAL GaussianModel.construct_list_of_attributes
(scene/gaussian_model.py:890-903), whose source is cut out with `ast` and executed on a stand-in object with the
reference's tensor shapes (the module itself cannot be imported here: plyfile / FrEIA are absent)."""
import ast
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scene/gaussian_model.py"
text = open(SRC).read()
fn = None
for node in ast.walk(ast.parse(text)):
    if isinstance(node, ast.FunctionDef) and node.name == "construct_list_of_attributes":
        fn = ast.get_source_segment(text, node)
ns = {}
exec("import textwrap\n", ns)
exec(compile(__import__("textwrap").dedent(fn), SRC, "exec"), ns)


class Stand:
    pass


def attrs(K):
    s = Stand()
    P = 2
    s._features_dc, s._features_rest = torch.zeros(P, 1, 3), torch.zeros(P, K, 3)
    s._scaling, s._rotation = torch.zeros(P, 3), torch.zeros(P, 4)
    return ns["construct_list_of_attributes"](s)


torch.save({"attributes_sh3": attrs(15), "attributes_sh0": attrs(0)}, os.path.join(HERE, "ply_golden.pt"))
print(len(attrs(15)), attrs(15)[:8], attrs(15)[-8:])
