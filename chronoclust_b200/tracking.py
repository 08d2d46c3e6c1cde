"""Cluster tracking across timepoints -- host code (BASELINE.json north_star keeps tracking on the host).

Behavioural restatement of tracking/cluster_tracker.py:7-148 of the reference: lineage labels (A, B, A|1,
(A,B), ...) from shared pcore ids, and historical association by nearest previous pcore MC.  The inner
distance scan of the historical tracker is SURVEY 8f-1 ("next" row): from 65 536 (current, previous) pcore pairs on
it runs on the device (ccb_assoc_nearest), below that it is the host loop; both give the same associations.
"""
import string
from collections import defaultdict, deque


class TrackByLineage(object):
    def __init__(self):
        self.child_clusters = []
        self.parent_clusters = []
        self.letters = deque(string.ascii_uppercase)
        self.split_per_id = {letter: 0 for letter in self.letters}

    def repopulate_letters(self, iteration):
        self.letters = deque(letter * iteration for letter in string.ascii_uppercase)

    def get_new_letter(self):
        letter = self.letters.popleft()
        if not self.letters:  # Z used up -> AA..ZZ, then AAA.. (cluster_tracker.py:20-24)
            self.repopulate_letters(len(letter) + 1)
        return letter

    def add_new_child_cluster(self, cluster):
        self.child_clusters.append(cluster)

    def get_parent_pcore_to_id(self):
        out = {}
        for parent in self.parent_clusters:
            for pcore in parent.pcore_ids:
                out[pcore] = parent.id
        return out

    def calculate_ids(self):
        # presentation order: ascending (rounded) weight, stable (cluster_tracker.py:33-34)
        if None not in [c.cumulative_weight for c in self.child_clusters]:
            self.child_clusters.sort(key=lambda c: c.cumulative_weight)
        offspring = defaultdict(list)
        parent_of_pcore = self.get_parent_pcore_to_id()
        for cluster in self.child_clusters:
            cluster.set_parents(parent_pcores_to_id=parent_of_pcore)
            if len(cluster.parents) == 0:  # made of new pcore MCs only -> new letter
                cluster.add_parent(id=self.get_new_letter())
            for parent in cluster.get_parents():
                offspring[parent].append(cluster)
        for parent, children in offspring.items():
            # the child holding most pcore MCs keeps the parent's label; the others are splits parent|n,
            # numbered on from the splits that parent has already had
            children = sorted(children, key=lambda c: len(c.pcore_ids), reverse=True)
            nsplit = self.split_per_id.get(parent, 0)
            for rank, child in enumerate(children):
                if rank == 0:
                    child.add_id(parent)
                else:
                    nsplit += 1
                    child.add_id(f'{parent}|{nsplit}')
            self.split_per_id[parent] = nsplit
        self.assign_child_id()

    def transfer_child_to_parent(self):
        self.parent_clusters = self.child_clusters
        self.child_clusters = []

    def assign_child_id(self):
        """One label -> itself; several labels -> a merge, nested as ((A,B),C) in sorted order."""
        for child in self.child_clusters:
            labels = sorted(child.id)
            if len(labels) == 1:
                child.id = labels[0]
            else:
                merged = f'({labels[0]},{labels[1]})'
                for extra in labels[2:]:
                    merged = f'({merged},{extra})'
                child.id = merged


class TrackByHistoricalAssociation(object):
    # the association scan moves to the device (ccb_assoc_nearest) from this many (current, previous) pcore pairs
    # on; None keeps the host loop.  Below it the two host <-> device copies cost more than the loop.
    device_scan_min_pairs = 1 << 16

    device = 0  # CUDA ordinal of the association scan (app.run sets it to its own `device`)

    def __init__(self):
        self.current_clusters = []
        self.previous_timepoint_clusters = []

    def set_current_clusters(self, clusters):
        self.current_clusters = clusters

    def track_cluster_history(self):
        if len(self.previous_timepoint_clusters) == 0:
            for cluster in self.current_clusters:
                cluster.add_historical_associate(None)
            return
        cur = [(cluster, pcore) for cluster in self.current_clusters for pcore in cluster.pcore_objects]
        prev = [(c.id, p) for c in self.previous_timepoint_clusters for p in c.pcore_objects]
        if self.device_scan_min_pairs is not None and prev and len(cur) * len(prev) >= self.device_scan_min_pairs \
                and self._cuda_ready():
            best = self._device_scan([p for _, p in cur], [p for _, p in prev], self.device)
            for (cluster, _), j in zip(cur, best):
                cluster.add_historical_associate(prev[j][0])
                cluster.add_historical_associate_pcore(prev[j][1].id)
            return
        for cluster, pcore in cur:
            best_d, best_cluster, best_pcore = None, None, None
            for prev_id, prev_pcore in prev:
                d = pcore.get_projected_dist_to_point(prev_pcore.cluster_centroids)
                if best_d is None or d < best_d:  # strict <: first wins
                    best_d, best_cluster, best_pcore = d, prev_id, prev_pcore.id
            cluster.add_historical_associate(best_cluster)
            cluster.add_historical_associate_pcore(best_pcore)

    @staticmethod
    def _cuda_ready():
        """The device scan needs torch (device buffers) and a CUDA device; without them the host loop below runs."""
        try:
            import torch

            return torch.cuda.is_available()
        except ImportError:
            return False

    @staticmethod
    def _device_scan(cur_pcores, prev_pcores, device=0):
        """the O(P_cur * P_prev * D) scan on the device (ccb_assoc_nearest, kernel k_assoc) -- the same
        argmin, bit for bit (sequential sum over d, the query's stored centroid / preference vector, first wins)."""
        import numpy as np
        import torch

        from . import _lib

        D = len(cur_pcores[0].cluster_centroids)
        ccen = np.ascontiguousarray([p.cluster_centroids for p in cur_pcores], np.float64).reshape(len(cur_pcores), D)
        pref = np.ascontiguousarray([p.preferred_dimension_vector for p in cur_pcores], np.float64).reshape(len(cur_pcores), D)
        pcen = np.ascontiguousarray([p.cluster_centroids for p in prev_pcores], np.float64).reshape(len(prev_pcores), D)
        ks = np.unique(pref[pref != 1.0])
        if len(ks) > 1:
            raise ValueError("preference vectors with more than one weight value")
        k = float(ks[0]) if len(ks) else 1.0
        mask = ((pref != 1.0) * (np.uint64(1) << np.arange(D, dtype=np.uint64))).sum(axis=1).astype(np.uint64)
        dev = torch.device("cuda", device)
        tc, tp = torch.from_numpy(ccen).to(dev), torch.from_numpy(pcen).to(dev)
        tm = torch.from_numpy(mask.view(np.int64)).to(dev)
        best = torch.empty(len(cur_pcores), dtype=torch.int32, device=dev)
        dist = torch.empty(len(cur_pcores), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().ccb_assoc_nearest(device, None, tc.data_ptr(), tm.data_ptr(), len(cur_pcores), tp.data_ptr(),
                                                len(prev_pcores), D, k, best.data_ptr(), dist.data_ptr()))
        torch.cuda.synchronize(dev)
        return best.cpu().numpy().tolist()

    def transfer_current_to_previous(self):
        self.previous_timepoint_clusters = self.current_clusters
        self.current_clusters = []
