"""Host-side mirrors of the reference's state objects, backed by arrays exported from the device.

Microcluster  objects/microcluster.py:18-260 -- same attribute names; `points` is materialised lazily
              from the per-row assignment array instead of being filled one Python list at a time.
FinalCluster  the merged PredeconMC the offline phase emits (predecon.py:69-84): id set, CF1, CF2,
              cumulative_weight, cluster_centroids, preferred_dimension_vector.
Cluster       objects/cluster.py:5-115 -- the tracking DTO (host code; tracking stays on the host).
"""
import copy as _copy

import numpy as np


def _seq_sum(values):
    s = 0.0
    for v in values:  # sequential, like numba's np.sum
        s += float(v)
    return s


class Microcluster(object):
    def __init__(self, cf1, cf2, id=None, cumulative_weight=0, preferred_dimension_vector=None,
                 cluster_centroids=None, creation_time_in_hrs=0, prev_outlier_id=None, points_source=None):
        self.id = set() if id is None else id
        self.CF1 = cf1
        self.CF2 = cf2
        self.cumulative_weight = cumulative_weight
        self.preferred_dimension_vector = preferred_dimension_vector
        self.cluster_centroids = cluster_centroids
        self.creation_time_in_hrs = creation_time_in_hrs
        self.prev_pcore_id = None
        self.prev_outlier_id = prev_outlier_id
        self._points = None
        self._points_source = points_source  # callable uid -> dict(row -> list)

    # -- lazily materialised {row index: coordinates} of the current timepoint, insertion = row order
    @property
    def points(self):
        if self._points is None:
            src = self._points_source
            self._points = src(self.prev_outlier_id) if src is not None else {}
        return self._points

    @points.setter
    def points(self, value):
        self._points = value

    def reset_points(self):
        self._points = {}

    def __getstate__(self):
        # the per-cell dictionary is only pickled if somebody already materialised it (the reference
        # pickles every cell of the timepoint; the next timepoint resets it anyway, hddstream.py:208-213)
        st = dict(self.__dict__)
        st["_points"] = self._points if self._points is not None else {}
        st["_points_source"] = None
        return st

    def __deepcopy__(self, memo):
        new = Microcluster.__new__(Microcluster)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_points_source":
                new.__dict__[k] = v  # shared read-only view provider; keeps the copy lazy
            else:
                new.__dict__[k] = _copy.deepcopy(v, memo)
        return new

    def update_prev_outlier_id(self, outlier_id):
        self.prev_outlier_id = outlier_id

    def update_prev_pcore_id(self, pcore_id):
        self.prev_pcore_id = pcore_id

    # -- maths (host restatements, sequential over dimensions) ------------------------------------
    def set_centroid(self):
        self.cluster_centroids = np.asarray(self.CF1, dtype=np.float64) / self.cumulative_weight

    def update_preferred_dimensions(self, variance_threshold_squared, k_constant):
        w = self.cumulative_weight
        cf1, cf2 = np.asarray(self.CF1, np.float64), np.asarray(self.CF2, np.float64)
        var = cf2 / w - np.square(cf1 / w)
        self.preferred_dimension_vector = np.array(
            [k_constant if s <= variance_threshold_squared else 1.0 for s in var])

    def add_new_point(self, new_point_values, new_point_timestamp, new_point_idx, new_point_weight=1,
                      update_centroid=True):
        x = np.asarray(new_point_values, np.float64)
        self.CF1 = np.add(self.CF1, x)
        self.CF2 = np.add(self.CF2, np.square(x))
        self.cumulative_weight += new_point_weight
        self.points[new_point_idx] = x.tolist()
        if update_centroid:
            self.set_centroid()

    def get_projected_dist_to_point(self, other_point):
        c = np.asarray(self.cluster_centroids, np.float64)
        p = np.asarray(self.preferred_dimension_vector, np.float64)
        x = np.asarray(other_point, np.float64)
        return _seq_sum(np.square(x - c) / p)

    def calculate_projected_radius_squared(self):
        w = self.cumulative_weight
        cf1, cf2 = np.asarray(self.CF1, np.float64), np.asarray(self.CF2, np.float64)
        p = np.asarray(self.preferred_dimension_vector, np.float64)
        return _seq_sum((cf2 / w - np.square(cf1 / w)) / p)

    def get_copy(self):
        return Microcluster(cf1=np.copy(self.CF1), cf2=np.copy(self.CF2), cumulative_weight=self.cumulative_weight)

    def get_copy_with_new_point(self, datapoint, variance_threshold_squared, k_constant):
        tmp = self.get_copy()
        tmp.add_new_point(datapoint, -1, -1)
        tmp.update_preferred_dimensions(variance_threshold_squared, k_constant)
        return tmp

    def is_core(self, radius_threshold_squared, density_threshold, max_subspace_dimensionality):
        r2 = self.calculate_projected_radius_squared()
        cnt = int((np.asarray(self.preferred_dimension_vector) > 1).sum())
        return bool(r2 <= radius_threshold_squared and self.cumulative_weight >= density_threshold
                    and cnt <= max_subspace_dimensionality)


class FinalCluster(Microcluster):
    """What HDDStream.final_clusters holds: the cluster PreDeCon merged out of pcore MCs."""

    def __init__(self, member_ids_in_claim_order, cf1, cf2, weight, centroid, pref):
        ids = set()
        for m in member_ids_in_claim_order:  # claim order -> CPython set layout -> printed pcore_ids order
            ids.add(int(m))
        Microcluster.__init__(self, cf1=cf1, cf2=cf2, id=ids, cumulative_weight=weight,
                              preferred_dimension_vector=pref, cluster_centroids=centroid)
        self.centroid = centroid
        self.core_status = False
        self.members_in_claim_order = [int(m) for m in member_ids_in_claim_order]


class Cluster(object):
    """Tracking DTO (objects/cluster.py:5-115); plain host object."""

    def __init__(self, pcore_ids, cluster_centroid=None, cumulative_weight=None, preferred_dimensions=None):
        self.pcore_ids = pcore_ids
        self.id = set()
        self.parents = set()
        self.centroid = cluster_centroid
        self.cumulative_weight = cumulative_weight
        self.pcore_objects = []
        self.historical_associates = set()
        self.preferred_dimensions = preferred_dimensions
        self.historical_associates_pcores = set()

    def add_id(self, id):
        self.id.update([id])

    def add_parent(self, id):
        self.parents.update([id])

    def set_parents(self, parent_pcores_to_id):
        for pcore in self.pcore_ids:
            if pcore in parent_pcores_to_id:
                self.add_parent(parent_pcores_to_id[pcore])

    def get_parents(self):
        return self.parents

    def add_pcore_objects(self, pcore_id_to_object):
        for pcore_id in self.pcore_ids:
            self.pcore_objects.append(_copy.deepcopy(pcore_id_to_object[pcore_id]))

    def add_historical_associate(self, associate):
        self.historical_associates.update([associate])

    def add_historical_associate_pcore(self, pcore_id):
        self.historical_associates_pcores.update(pcore_id)

    def get_historical_associates_as_str(self):
        return '&'.join(str(s) for s in sorted(self.historical_associates))

    def get_historical_associates_pcore_as_str(self):
        return '&'.join(str(s) for s in self.historical_associates_pcores)

    def get_preferred_dimensions_as_str(self):
        return ';'.join(str(s) for s in self.preferred_dimensions)

    def get_pcore_ids_as_str(self):
        return '|'.join(str(s) for s in self.pcore_ids)

    def get_projected_dist_to_point(self, other_point):
        dist = 0.0
        for c_i, p_i, d_i in zip(self.centroid, other_point, self.preferred_dimensions):
            dist += ((float(p_i) - float(c_i)) ** 2) / float(d_i)
        return dist

    def get_dist_to_point(self, other_point):
        dist = 0.0
        for i, c in enumerate(self.centroid):
            dist += (float(other_point[i]) - float(c)) ** 2
        return dist
