"""Joint min-max normalisation of all timepoints (host component; north_star keeps scaling on the host).

Drop-in for the reference's `Scaler` (scaling/scaler.py:11-53): same constructor and method names, and the same
sklearn `MinMaxScaler` underneath, so the scaled fp64 values that feed the bit-exact path are identical.  Differences
in HOW it gets there:

* the fit streams the timepoint files one at a time through `partial_fit` (running column minima / maxima are exact, so
  the fitted `scale_` / `min_` equal those of one fit over the concatenation, bit for bit) instead of holding every cell
  of every timepoint in one Python list;
* `device_vectors()` hands the fitted `(scale_, min_)` to `HDDStream.online_microcluster_maintenance(..., scaler=)`, which
  applies the transform on the device behind the host -> device copy (`ccb_ingest_scaled`);
* `fit_columns_on_device` computes a file's column minima / maxima with `ccb_colminmax` (np.nanmin / np.nanmax semantics)
  for callers whose data already lives on the GPU.
"""
import numpy as np
import pandas as pd
from sklearn.preprocessing import MinMaxScaler


def _read_cells(path):
    return pd.read_csv(path, header=0, sep=',').to_numpy()


class Scaler(object):
    def __init__(self, data_files=None):
        self.scaler = MinMaxScaler()
        self.input_data = []
        self.files = list(data_files) if data_files is not None else []
        for path in self.files:
            self.scaler.partial_fit(_read_cells(path))

    # ---- reference API ---------------------------------------------------------------------------------------------
    def fit_scaler(self, data):
        """Refit on an in-memory array (scaler.py:38-41); the array is remembered as the reference does."""
        self.scaler = MinMaxScaler().fit(data)
        self.set_input_data(data)

    def scale_data(self, data):
        return self.scaler.transform(data)

    def reverse_scaling(self, data):
        return self.scaler.inverse_transform(data)

    def set_input_data(self, data):
        self.input_data = data

    def get_input_data(self):
        """The cells the scaler was set up with.  After a file-streamed fit they are re-read on demand."""
        if len(self.input_data) == 0 and self.files:
            self.input_data = np.concatenate([_read_cells(p) for p in self.files], axis=0)
        return self.input_data

    # ---- device side -----------------------------------------------------------------------------------
    def device_vectors(self):
        """(scale_, min_) as contiguous fp64 vectors for ccb_ingest_scaled: x * scale_ + min_, two roundings."""
        return (np.ascontiguousarray(self.scaler.scale_, np.float64), np.ascontiguousarray(self.scaler.min_, np.float64))

    @staticmethod
    def fit_columns_on_device(x_dev, device=0):
        """Column minima / maxima of a CUDA tensor [N, D] (fp64, row-major) ignoring NaN -- the reductions of
        MinMaxScaler.partial_fit -- through ccb_colminmax.  Returns two numpy vectors."""
        import torch

        from . import _lib

        n, d = x_dev.shape
        mn = torch.empty(d, dtype=torch.float64, device=x_dev.device)
        mx = torch.empty(d, dtype=torch.float64, device=x_dev.device)
        _lib.check(_lib.lib().ccb_colminmax(device, None, x_dev.data_ptr(), n, x_dev.stride(0), d, mn.data_ptr(),
                                            mx.data_ptr()))
        torch.cuda.synchronize(x_dev.device)
        return mn.cpu().numpy(), mx.cpu().numpy()
