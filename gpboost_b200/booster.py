"""Host-side mirror of the reference's `Dataset` / `Booster` (python-package/gpboost/basic.py:1040-2322, 2323-4170)
for the hot path: dense numerical matrices, `objective=regression`, optional GPModel (GPBoost algorithm).
Same C entry points (`LGBM_*`), bound with ctypes; `_lib` selects the shared library like in GPModel."""
import ctypes

import numpy as np

from .basic import GPBoostError, c_str, _dptr
from .libpath import load_lib

C_API_DTYPE_FLOAT32, C_API_DTYPE_FLOAT64 = 0, 1
C_API_PREDICT_NORMAL, C_API_PREDICT_RAW_SCORE = 0, 1


def param_dict_to_str(params):
    return " ".join("%s=%s" % (k, str(v).lower() if isinstance(v, bool) else v) for k, v in (params or {}).items())


class Dataset(object):
    def __init__(self, data, label=None, params=None, _lib=None):
        self._LIB = load_lib() if _lib is None else _lib
        data = np.asarray(data)
        # float32 matrices are passed as they are (C_API_DTYPE_FLOAT32), like the reference's package (basic.py: __init_from_np2d)
        dt = np.float32 if data.dtype == np.float32 else np.float64
        data = np.ascontiguousarray(data, dtype=dt)
        if data.ndim != 2:
            raise ValueError("'data' needs to be a 2-D array")
        self.num_data, self.num_feature = data.shape
        self.handle = ctypes.c_void_p()
        self._safe_call(self._LIB.LGBM_DatasetCreateFromMat(
            data.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(C_API_DTYPE_FLOAT32 if dt == np.float32 else C_API_DTYPE_FLOAT64), ctypes.c_int32(self.num_data),
            ctypes.c_int32(self.num_feature), ctypes.c_int(1), c_str(param_dict_to_str(params)), None, ctypes.byref(self.handle)))
        if label is not None:
            self.set_label(label)

    def _safe_call(self, ret):
        if ret != 0:
            raise GPBoostError(self._LIB.LGBM_GetLastError().decode("utf-8"))

    def set_label(self, label):
        lab = np.ascontiguousarray(np.asarray(label, dtype=np.float32).reshape(-1))  # label_t = float (meta.h:49)
        if lab.shape[0] != self.num_data:
            raise ValueError("Length of label is not same with #data")
        self.label = lab
        self._safe_call(self._LIB.LGBM_DatasetSetField(self.handle, c_str("label"), lab.ctypes.data_as(ctypes.c_void_p),
                                                       ctypes.c_int(lab.shape[0]), ctypes.c_int(C_API_DTYPE_FLOAT32)))

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value is not None:
                self._LIB.LGBM_DatasetFree(self.handle)
                self.handle = None
        except Exception:
            pass


class Booster(object):
    def __init__(self, params=None, train_set=None, gp_model=None, model_str=None, model_file=None, _lib=None):
        self._LIB = load_lib() if _lib is None else _lib
        self.train_set, self.gp_model = train_set, gp_model
        self.handle = ctypes.c_void_p()
        if model_str is not None or model_file is not None:  # prediction-only booster (Booster(model_file=...), basic.py:2404-2425)
            n_it = ctypes.c_int(0)
            if model_str is not None:
                self._safe_call(self._LIB.LGBM_BoosterLoadModelFromString(c_str(model_str), ctypes.byref(n_it), ctypes.byref(self.handle)))
            else:
                self._safe_call(self._LIB.LGBM_BoosterCreateFromModelfile(c_str(model_file), ctypes.byref(n_it), ctypes.byref(self.handle)))
            return
        params = dict(params or {})
        if gp_model is not None:
            params["has_gp_model"] = True
            self._safe_call(self._LIB.LGBM_GPBoosterCreate(train_set.handle, c_str(param_dict_to_str(params)), gp_model.handle,
                                                           ctypes.byref(self.handle)))
        else:
            self._safe_call(self._LIB.LGBM_BoosterCreate(train_set.handle, c_str(param_dict_to_str(params)), ctypes.byref(self.handle)))

    def _safe_call(self, ret):
        if ret != 0:
            raise GPBoostError(self._LIB.LGBM_GetLastError().decode("utf-8"))

    def update(self):
        """One boosting iteration (Booster.update, basic.py:2846-2905). Returns True when no further split was possible."""
        fin = ctypes.c_int(0)
        self._safe_call(self._LIB.LGBM_BoosterUpdateOneIter(self.handle, ctypes.byref(fin)))
        return fin.value == 1

    def current_iteration(self):
        it = ctypes.c_int(0)
        self._safe_call(self._LIB.LGBM_BoosterGetCurrentIteration(self.handle, ctypes.byref(it)))
        return it.value

    def model_to_string(self, start_iteration=0, num_iteration=-1):
        n = ctypes.c_int64(0)
        buf_len = 1 << 20
        buf = ctypes.create_string_buffer(buf_len)
        self._safe_call(self._LIB.LGBM_BoosterSaveModelToString(self.handle, ctypes.c_int(start_iteration), ctypes.c_int(num_iteration), ctypes.c_int(0),
                                                                ctypes.c_int64(buf_len), ctypes.byref(n), buf))
        if n.value > buf_len:
            buf_len = n.value
            buf = ctypes.create_string_buffer(buf_len)
            self._safe_call(self._LIB.LGBM_BoosterSaveModelToString(self.handle, ctypes.c_int(start_iteration), ctypes.c_int(num_iteration), ctypes.c_int(0),
                                                                    ctypes.c_int64(buf_len), ctypes.byref(n), buf))
        return buf.value.decode("utf-8")

    def save_model(self, filename):
        self._safe_call(self._LIB.LGBM_BoosterSaveModel(self.handle, ctypes.c_int(0), ctypes.c_int(-1), ctypes.c_int(0), c_str(filename)))
        return self

    def inner_predict_train(self):
        """Raw training scores F (Booster.__inner_predict(0), basic.py:3964)."""
        n = ctypes.c_int64(0)
        self._safe_call(self._LIB.LGBM_BoosterGetNumPredict(self.handle, ctypes.c_int(0), ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        self._safe_call(self._LIB.LGBM_BoosterGetPredict(self.handle, ctypes.c_int(0), ctypes.byref(n), _dptr(out)))
        return out

    def predict(self, data, raw_score=True, start_iteration=0, num_iteration=-1):
        data = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
        n = ctypes.c_int64(0)
        out = np.empty(data.shape[0], dtype=np.float64)
        self._safe_call(self._LIB.LGBM_BoosterPredictForMat(
            self.handle, data.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(C_API_DTYPE_FLOAT64), ctypes.c_int32(data.shape[0]),
            ctypes.c_int32(data.shape[1]), ctypes.c_int(1), ctypes.c_int(C_API_PREDICT_RAW_SCORE if raw_score else C_API_PREDICT_NORMAL),
            ctypes.c_int(start_iteration), ctypes.c_int(num_iteration), c_str(""), ctypes.byref(n), _dptr(out)))
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value is not None:
                self._LIB.LGBM_BoosterFree(self.handle)
                self.handle = None
        except Exception:
            pass


def parse_model_string(s):
    """Trees of a LightGBM/GPBoost text model -> list of dicts of numpy arrays (io/gbdt_model_text.cpp, Tree::ToString)."""
    trees = []
    cur = None
    for line in s.splitlines():
        if line.startswith("Tree="):
            cur = {}
            trees.append(cur)
        elif cur is not None and "=" in line:
            k, v = line.split("=", 1)
            if k in ("split_feature", "left_child", "right_child", "leaf_count", "internal_count", "decision_type"):
                cur[k] = np.array([int(x) for x in v.split()], dtype=np.int64)
            elif k in ("threshold", "leaf_value", "split_gain", "internal_value", "leaf_weight", "internal_weight"):
                cur[k] = np.array([float(x) for x in v.split()], dtype=np.float64)
            elif k in ("num_leaves", "num_cat"):
                cur[k] = int(v)
            elif k == "shrinkage":
                cur[k] = float(v)
        elif line.startswith("end of trees"):
            break
    return trees
