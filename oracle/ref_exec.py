"""Run the reference's OWN source for the measurement / export path (TEST INFRASTRUCTURE).

The reference scripts cannot be imported: they execute at import, with hard-coded paths,
trained weights and Detectron2 (nn_inference.py:168, :225, :309, :482).  But the functions and
loops of the hot path are plain Python over module globals, so this module takes their AST nodes
out of ``/root/reference/nn_inference.py`` (and ``backup_main.py`` as the second witness),
compiles them UNMODIFIED and executes them in a namespace whose globals are stubs for what the
image does not have:

  node                                   reference lines        how it is run
  ``rle_decode``                         nn_inference.py:237-251   called directly
  ``rle_encoding``                       :253-263                  called directly
  ``postprocess_masks``                  :265-306                  called directly
  export loop ``for name in images_name``:313-336                  exec'd with ``cv2.imread`` /
                                                                   ``predictor`` stubs, csv read back
  ``midpoint``                           :339-340                  called by GetMask_Contours
  ``GetCounts``                          :355-366                  called directly / by the driver loop
  ``GetMask_Contours``                   :371-459                  called directly / by the driver loop
  driver loop ``for k in keywds``        :487-570                  exec'd once per keyword (see
                                                                   ``run_class_driver``)
  ``GetMask_Contours`` (no class filter) backup_main.py:429-497    called directly

Stubs (everything else is the real library as installed here -- numpy, cv2, scipy, pandas, csv):
  * ``predictor(im)`` returns ``{"instances": <oracle d2 Instances after detector_postprocess>}``
    -- Detectron2 is not installable; oracle/d2.py restates its glue and the paste itself is torch's
    own ``grid_sample``;
  * ``imutils`` / ``contours`` / ``perspective``: oracle/imutils_port.py (imutils 0.5.4 is absent);
  * ``erosion`` / ``dilation`` / ``label`` (scikit-image is absent): the SciPy routines
    scikit-image wraps, see oracle/cleanup.py -- THIS ONE SUBSTITUTION REMAINS;
  * ``plt.figure``, ``Image.fromarray(...).save``, ``cv2.imwrite``, ``Visualizer`` : no-ops
    (they produce no data); ``cv2.imread`` / ``os.listdir`` serve the in-memory fixture images.

Nothing here is copied from the reference: the nodes are read from where they lie at run time,
which is why this module only works where ``/root/reference`` exists (the build container);
``tests/golden/make_ref_golden.py`` freezes its outputs into ``tests/golden/ref_exec_*.npz`` for
the GPU box.
"""
from __future__ import annotations

import ast
import contextlib
import csv
import io
import os
import tempfile
import types
from typing import Any, Dict, List, Optional, Sequence, Tuple

import cv2 as _cv2
import numpy as np
import pandas as pd
import scipy.ndimage as ndi
from scipy.spatial import distance as _dist

from . import imutils_port

REF_DIR = os.environ.get("UWCV_REFERENCE_DIR", "/root/reference")
NN_INFERENCE = os.path.join(REF_DIR, "nn_inference.py")
BACKUP_MAIN = os.path.join(REF_DIR, "backup_main.py")

LIST_NAMES = ["lengthList", "widthList", "circularEDList", "aspectRatioList", "circularityList",
              "chordsList", "ferretList", "roundList", "sphereList"]
COUNT_LISTS = ["SList", "WTList", "PTList", "PList"]
# CSV column order of the reference (:561, :569) expressed as list names
CSV_ORDER = ["ferretList", "aspectRatioList", "roundList", "circularityList", "sphereList",
             "lengthList", "widthList", "circularEDList", "chordsList"]


def available() -> bool:
    return os.path.isfile(NN_INFERENCE)


# ------------------------------------------------------------------------------------------
# AST extraction
# ------------------------------------------------------------------------------------------

def _parse(path: str) -> ast.Module:
    with open(path, "r") as f:
        return ast.parse(f.read(), filename=path)


def _function(tree: ast.Module, name: str) -> ast.FunctionDef:
    hits = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name]
    if not hits:
        raise LookupError(f"reference has no top-level function {name!r}")
    return hits[-1]            # a later def shadows an earlier one, as at import


def _toplevel_for(tree: ast.Module, target: str, iter_name: str) -> ast.For:
    for n in tree.body:
        if (isinstance(n, ast.For) and isinstance(n.target, ast.Name) and n.target.id == target
                and isinstance(n.iter, ast.Name) and n.iter.id == iter_name):
            return n
    raise LookupError(f"reference has no top-level loop 'for {target} in {iter_name}'")


def _toplevel_assign(tree: ast.Module, name: str, after_line: int = 0) -> ast.Assign:
    for n in tree.body:
        if (isinstance(n, ast.Assign) and n.lineno > after_line and len(n.targets) == 1
                and isinstance(n.targets[0], ast.Name) and n.targets[0].id == name):
            return n
    raise LookupError(f"reference has no top-level assignment to {name!r}")


def _toplevel_expr_after(tree: ast.Module, line: int) -> ast.stmt:
    for n in tree.body:
        if n.lineno > line:
            return n
    raise LookupError("nothing after line %d" % line)


def _compile(nodes: Sequence[ast.stmt], path: str):
    mod = ast.Module(body=list(nodes), type_ignores=[])
    return compile(mod, path, "exec")


def node_lines(path: str = NN_INFERENCE) -> Dict[str, Tuple[int, int]]:
    """(first, last) source line of every node this module executes -- written into the golden
    files so a citation can be checked against the reference."""
    t = _parse(path)
    out = {}
    for name in ("rle_decode", "rle_encoding", "postprocess_masks", "midpoint", "GetCounts",
                 "GetMask_Contours"):
        try:
            f = _function(t, name)
            out[name] = (f.lineno, f.end_lineno)
        except LookupError:
            pass
    for key, tgt, it in (("export_loop", "name", "images_name"), ("class_driver", "k", "keywds")):
        try:
            f = _toplevel_for(t, tgt, it)
            out[key] = (f.lineno, f.end_lineno)
        except LookupError:
            pass
    return out


# ------------------------------------------------------------------------------------------
# stubs
# ------------------------------------------------------------------------------------------

class _Cv2Proxy:
    """cv2 with the two file-system calls replaced; every other attribute is the real cv2."""

    def __init__(self, runner: "ReferenceRunner"):
        self._r = runner

    def __getattr__(self, name):
        return getattr(_cv2, name)

    def imwrite(self, *a, **k):
        return True

    def imread(self, path, *a, **k):
        name = os.path.basename(path)
        self._r.current = name
        return self._r.images[name][0]


class _OsProxy:
    def __init__(self, runner: "ReferenceRunner"):
        self._r = runner
        self.path = os.path

    def __getattr__(self, name):
        return getattr(os, name)

    def listdir(self, path):
        return list(self._r.images.keys())


class _NoOp:
    def __getattr__(self, name):
        return _NoOp()

    def __call__(self, *a, **k):
        return _NoOp()


_CROSS = ndi.generate_binary_structure(2, 1)
_FULL = np.ones((3, 3), dtype=bool)


def _sk_dilation(mask):
    """skimage.morphology.dilation, default footprint (3 x 3 cross), default reflected border."""
    return ndi.grey_dilation(mask, footprint=_CROSS)


def _sk_erosion(mask):
    return ndi.grey_erosion(mask, footprint=_CROSS)


def _sk_label(mask):
    """skimage.measure.label, default connectivity (8 in 2-D)."""
    return ndi.label(mask, structure=_FULL)[0]


class _OneKeyword(list):
    """The ``keywds`` list of the driver loop (:485) made to iterate over ONE keyword:
    the loop as written dies on ``keywds[k]`` with ``k == 9`` after its first keyword (:570) and
    never resets the nine lists between keywords (:463-471), so each keyword is run as the
    loop's first (and only) iteration.  ``.index`` and everything else is the plain list."""

    def __init__(self, items, only):
        super().__init__(items)
        self._only = only

    def __iter__(self):
        return iter([self._only])


# ------------------------------------------------------------------------------------------
# runner
# ------------------------------------------------------------------------------------------

class ReferenceRunner:
    """Executes reference nodes.  ``images`` maps a file name to ``(im HxWx3 u8, instances)``
    where ``instances`` is what ``predictor(im)["instances"]`` holds (post-processed oracle
    Instances: ``pred_masks`` N x H x W bool, ``pred_classes``, ``scores``)."""

    def __init__(self, images: Optional[Dict[str, Tuple[np.ndarray, Any]]] = None,
                 path: str = NN_INFERENCE):
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        self.path = path
        self.tree = _parse(path)
        self.images = dict(images or {})
        self.current: Optional[str] = None
        self.ns: Dict[str, Any] = {}
        self._fresh_namespace()

    # -- namespace ---------------------------------------------------------------------
    def _predictor(self, im):
        name = self.current
        if name is None or self.images[name][0] is not im:
            for k, (img, _) in self.images.items():
                if img is im:
                    name = k
                    break
        return {"instances": self.images[name][1]}

    def _fresh_namespace(self) -> None:
        ns: Dict[str, Any] = {"__name__": "reference_exec", "__builtins__": __builtins__}
        ns.update(np=np, cv2=_Cv2Proxy(self), os=_OsProxy(self), pd=pd, csv=csv, dist=_dist,
                  imutils=types.SimpleNamespace(grab_contours=imutils_port.grab_contours,
                                                is_cv2=lambda: False),
                  contours=types.SimpleNamespace(sort_contours=imutils_port.sort_contours),
                  perspective=types.SimpleNamespace(order_points=imutils_port.order_points),
                  plt=_NoOp(), Image=_NoOp(), Visualizer=_NoOp(), ColorMode=_NoOp(),
                  multiclass_test_metadata=None,
                  binary_fill_holes=ndi.binary_fill_holes, erosion=_sk_erosion,
                  dilation=_sk_dilation, label=_sk_label, predictor=self._predictor)
        for n in LIST_NAMES + COUNT_LISTS:
            ns[n] = list()
        # DList / BList: the count lists of backup_main.py's GetCounts
        for n in ("DList", "BList"):
            ns[n] = list()
        funcs = []
        for name in ("rle_decode", "rle_encoding", "postprocess_masks", "midpoint",
                     "GetInference", "GetCounts", "GetMask_Contours"):
            try:
                funcs.append(_function(self.tree, name))
            except LookupError:
                pass
        exec(_compile(funcs, self.path), ns)
        self.ns = ns

    def reset_lists(self) -> None:
        for n in LIST_NAMES + COUNT_LISTS + ["DList", "BList"]:
            self.ns[n] = list()

    def lists(self) -> Dict[str, list]:
        return {n: list(self.ns[n]) for n in LIST_NAMES}

    def rows(self) -> np.ndarray:
        """The nine lists as K x 9 float64 rows in the reference's CSV column order."""
        cols = [np.asarray(self.ns[n], dtype=np.float64) for n in CSV_ORDER]
        return np.stack(cols, axis=1).reshape(-1, 9) if len(cols[0]) else np.zeros((0, 9))

    def list_dtypes(self) -> Dict[str, str]:
        """numpy / Python type of the list entries (float32 vs float64 matters to :523-527)."""
        return {n: (type(self.ns[n][0]).__name__ if self.ns[n] else "") for n in CSV_ORDER}

    # -- direct calls -------------------------------------------------------------------
    def call(self, name: str, *args, **kwargs):
        with contextlib.redirect_stdout(io.StringIO()):
            return self.ns[name](*args, **kwargs)

    def get_mask_contours(self, name: str, classes_of_interest: Optional[Sequence[int]] = None):
        """GetMask_Contours on fixture image ``name``; returns the rows it appended."""
        self.reset_lists()
        im = self.images[name][0]
        self.current = name
        self.ns["im"] = im
        with contextlib.redirect_stdout(io.StringIO()), _in_tmpdir():
            if classes_of_interest is None:      # backup_main.py form: no arguments
                self.ns["GetMask_Contours"]()
            else:
                self.ns["GetMask_Contours"](im, classes_of_interest=list(classes_of_interest))
        return self.rows()

    def get_counts(self, name: str) -> Dict[str, int]:
        self.reset_lists()
        self.current = name
        self.ns["im"] = self.images[name][0]
        self.call("GetCounts")
        return {n: int(self.ns[n][0]) for n in COUNT_LISTS if self.ns[n]}

    # -- module-level loops ----------------------------------------------------------------
    def run_export_loop(self) -> Tuple[List[str], List[str], str]:
        """:313-336 -- ``Img_ID`` / ``EncodedPixels`` lists and the text of R50_flip_.csv."""
        loop = _toplevel_for(self.tree, "name", "images_name")
        pre = [_toplevel_assign(self.tree, n, after_line=_function(self.tree, "postprocess_masks").end_lineno)
               for n in ("Img_ID", "EncodedPixels", "num", "conv")]
        post = [n for n in self.tree.body if loop.end_lineno < n.lineno <= loop.end_lineno + 4
                and not isinstance(n, ast.FunctionDef)]
        self.ns["images_name"] = list(self.images.keys())
        self.ns["inpath"] = "/fixture/"
        with contextlib.redirect_stdout(io.StringIO()), _in_tmpdir() as d:
            os.makedirs(os.path.join(d, "output"))
            exec(_compile(pre + [loop] + post, self.path), self.ns)
            text = ""
            out = os.path.join(d, "output", "R50_flip_.csv")
            if os.path.exists(out):
                with open(out) as f:
                    text = f.read()
        return list(self.ns["Img_ID"]), list(self.ns["EncodedPixels"]), text

    def run_class_driver(self, keyword: str) -> Dict[str, Any]:
        """:485-570 for one keyword over all fixture images: the nine lists as appended, the count
        lists, the smoothed ``ShapeDescriptor.csv`` text (:556-559) and how the loop ended."""
        loop = _toplevel_for(self.tree, "k", "keywds")
        self.reset_lists()
        ns = self.ns
        ns.update(tS=0, tWT=0, tPT=0, tP=0, count=0, test_img_path="/fixture/", x_c=0)
        base = ast.literal_eval(_toplevel_assign(self.tree, "keywds").value)
        ns["keywds"] = _OneKeyword(base, keyword)
        ended = "completed"
        with contextlib.redirect_stdout(io.StringIO()), _in_tmpdir() as d:
            try:
                exec(_compile([loop], self.path), ns)
            except IndexError as e:            # keywds[k] with k == 9 (:570)
                ended = "IndexError: %s" % e
            except Exception as e:             # e.g. imutils sort_contours on no contours
                ended = "%s: %s" % (type(e).__name__, e)
            shape_csv = ""
            p = os.path.join(d, "ShapeDescriptor.csv")
            if os.path.exists(p):
                with open(p) as f:
                    shape_csv = f.read()
        return dict(rows=self.rows(), dtypes=self.list_dtypes(),
                    counts={n: [int(v) for v in ns[n]] for n in COUNT_LISTS},
                    totals=dict(tS=int(ns["tS"]), tWT=int(ns["tWT"]), tPT=int(ns["tPT"]),
                                tP=int(ns["tP"])),
                    shape_csv=shape_csv, ended=ended, count=int(ns["count"]))


@contextlib.contextmanager
def _in_tmpdir():
    """The reference writes 'predicted_masks.jpg', 'ShapeDescriptor.csv', ... into the cwd."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)
