"""`import yfv2` registers the hyphen-named directory `yolo-fastestv2_b200/` as package `yfv2_b200`
and puts it on sys.path so its drop-in mirrors (`model.detector`, `utils.utils`, ...) resolve."""
import importlib.util
import os
import sys

PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "yolo-fastestv2_b200")
if "yfv2_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location("yfv2_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["yfv2_b200"] = _mod
    _spec.loader.exec_module(_mod)
if PKG_DIR not in sys.path:
    sys.path.insert(0, PKG_DIR)
