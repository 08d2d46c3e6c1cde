"""Cluster tracking across timepoints -- host code (BASELINE.json north_star keeps tracking on the host).

Behavioural restatement of tracking/cluster_tracker.py:7-148 of the reference: lineage labels (A, B, A|1,
(A,B), ...) from shared pcore ids, and historical association by nearest previous pcore MC.  The inner
distance scan of the historical tracker is SURVEY 8f-1 ("next" row): here it is still the host loop.
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
        for cluster in self.current_clusters:
            for pcore in cluster.pcore_objects:
                best_d, best_cluster, best_pcore = None, None, None
                for prev in self.previous_timepoint_clusters:
                    for prev_pcore in prev.pcore_objects:
                        d = pcore.get_projected_dist_to_point(prev_pcore.cluster_centroids)
                        if best_d is None or d < best_d:  # strict <: first wins
                            best_d, best_cluster, best_pcore = d, prev.id, prev_pcore.id
                cluster.add_historical_associate(best_cluster)
                cluster.add_historical_associate_pcore(best_pcore)

    def transfer_current_to_previous(self):
        self.previous_timepoint_clusters = self.current_clusters
        self.current_clusters = []
