"""Joint min-max normalisation of all timepoints -- host code (north_star keeps scaling on the host).

Same contract as scaling/scaler.py:11-53 of the reference, built on the same sklearn MinMaxScaler so
that the scaled fp64 values that feed the bit-exact path are identical.
"""
import numpy as np
import pandas as pd
from sklearn.preprocessing import MinMaxScaler


class Scaler(object):
    def __init__(self, data_files=None):
        self.scaler = MinMaxScaler()
        self.input_data = []
        if data_files is not None:
            parts = [pd.read_csv(f, header=0, sep=',').to_numpy() for f in data_files]
            self.fit_scaler(np.concatenate(parts, axis=0))

    def fit_scaler(self, data):
        self.scaler.fit(data)
        self.set_input_data(data)

    def scale_data(self, data):
        return self.scaler.transform(data)

    def reverse_scaling(self, data):
        return self.scaler.inverse_transform(data)

    def set_input_data(self, data):
        self.input_data = data

    def get_input_data(self):
        return self.input_data
