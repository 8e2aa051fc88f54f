"""yolo-fastestv2_b200: B200 (sm_100a) implementation of the Yolo-FastestV2 hot path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI of libyfv2.so (include/yfv2.h)
  yfv2_engine.py   ctypes binding + plan cache
  model/, utils/   drop-in mirrors of the reference's import surface (model.detector.Detector,
                   utils.utils.handel_preds / non_max_suppression / load_datafile, ...): put this
                   directory first on PYTHONPATH and the reference's train.py / test.py /
                   evaluation.py import these instead of their own modules.

The directory name is not an importable identifier; `import yfv2` (repo root) registers it as the
package `yfv2_b200`.
"""
