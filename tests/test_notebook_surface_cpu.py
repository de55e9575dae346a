"""CPU, build container only (needs the reference's notebooks): every free name the 01/02/03/05 train notebooks use
resolves after their cell-1 imports are pointed at this repository's pyfiles/ (the drop-in boundary is that
import surface)."""
import ast
import builtins
import json
import os
import subprocess
import sys

import pytest

import ref_harness

NOTEBOOKS = ["01-train_Conventional_SingleGAN", "02-train_SingleGAN_soloD",
             "03-train_Style-Restricted_GAN_nopretraining", "05-train_Style-Restricted_GAN"]
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ref_harness.REF_ROOT, "notebook")),
                                reason="reference notebooks not present")

_PROBE = r'''
import sys, types, json
sys.path.insert(0, sys.argv[1])
for n in ("matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.transforms", "tqdm"):
    try:
        __import__(n)
    except Exception:
        sys.modules[n] = types.ModuleType(n)
ns = {}
exec("from util import *\nfrom dataset import *\nfrom model import *\nfrom util_notebook import *", ns)
print(json.dumps(sorted(k for k in ns if not k.startswith("__"))))
'''


def _free_names(path):
    cells = json.load(open(path))["cells"]
    defined, used = set(), set()
    for c in cells:
        if c["cell_type"] != "code":
            continue
        src = "\n".join(l for l in "".join(c["source"]).split("\n") if not l.strip().startswith(("%", "!")))
        try:
            tree = ast.parse(src)
        except SyntaxError:
            continue
        for n in ast.walk(tree):
            if isinstance(n, ast.Name):
                (defined if isinstance(n.ctx, (ast.Store, ast.Del)) else used).add(n.id)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                defined.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                defined.add(n.name)
            elif isinstance(n, ast.arg):
                defined.add(n.arg)
    return {u for u in used if u not in defined and not hasattr(builtins, u)}


def test_train_notebooks_resolve_every_name_against_this_surface():
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "style-restricted_gan_b200", "pyfiles")
    r = subprocess.run([sys.executable, "-c", _PROBE, here], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    exported = set(json.loads(r.stdout.strip().splitlines()[-1]))
    for k in ("SingleGenerator", "SingleDiscriminator_solo_multi", "Encoder", "SRGAN_training", "SingleGAN_training",
              "FaceDataset", "get_class_label", "get_output_and_plot", "get_target", "weights_init", "MinMax"):
        assert k in exported, k
    for nb in NOTEBOOKS:
        missing = sorted(_free_names(os.path.join(ref_harness.REF_ROOT, "notebook", nb + ".ipynb")) - exported)
        assert missing == [], (nb, missing)


_PROBE04 = r'''
import sys, types, json
sys.path.insert(0, sys.argv[1])
for n in ("matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.transforms", "tqdm"):
    try:
        __import__(n)
    except Exception:
        sys.modules[n] = types.ModuleType(n)
ns = {}
# cell 1 of notebook 04, verbatim, plus the modules behind the names it uses without importing them
exec("from util import image_from_output, cuda2numpy, cuda2cpu, weights_init, plot_confusion_matrix\n"
     "from dataset import get_class_label, FaceDataset\nfrom model import MinMax\n"
     "from model import Encoder_classifier\nfrom util import pickle_load\n"
     "import evaluation\nfrom util_notebook import Classifier_training, do_test", ns)
print(json.dumps(sorted(k for k in ns if not k.startswith("__")) + sorted(dir(ns["evaluation"]))))
'''


def test_notebook_04_and_06_imports_resolve():
    """Cell 1 of notebooks 04 / 06 (`from util import ...`, `from dataset import ...`, `from model import MinMax`,
    `from util import pickle_load`) and the f4 additions (Encoder_classifier, the evaluation module, the classifier job)."""
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "style-restricted_gan_b200", "pyfiles")
    r = subprocess.run([sys.executable, "-c", _PROBE04, here], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    names = set(json.loads(r.stdout.strip().splitlines()[-1]))
    for k in ("image_from_output", "cuda2numpy", "cuda2cpu", "weights_init", "plot_confusion_matrix", "get_class_label",
              "FaceDataset", "MinMax", "Encoder_classifier", "pickle_load", "Classifier_training", "do_test",
              "GAN_evaluation", "evaluation_init", "compute_prdc", "vgg_model"):
        assert k in names, k
