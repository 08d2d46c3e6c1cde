"""chronoclust_b200.app.run -- drop-in for chronoclust.app.run (reference app.py:32-226).

Same keyword parameters, same output files (result.csv, cluster_points_D{t}.csv, parameters.csv,
logs/Chronoclust.log, program_images/*).  The per-timepoint clustering (the hot path) runs on the GPU
behind chronoclust_b200.hddstream.HDDStream; scaling, lineage / association tracking and file I/O stay
on the host as BASELINE.json's north_star prescribes.  The per-cell output is assembled from the
device's per-row assignment array (vectorised) instead of per-cell Python lists.
"""
import csv
import logging
import os
import pickle
from collections import defaultdict
from decimal import ROUND_HALF_UP, Decimal

import numpy as np
import pandas as pd

from .hddstream import HDDStream
from .objects import Cluster
from .scaling import Scaler
from .tracking import TrackByHistoricalAssociation, TrackByLineage

HDDSTREAM_OBJ = 'hddstream'
TRACKER_HISTORICAL_ASSOC = 'tracking_by_historical_association'
TRACKER_LINEAGE = 'tracking_by_lineage'


def run(data, output_directory, gating_centroid_file=None, normalise_data=True, restore_program=False,
        param_beta=0.8, param_delta=0.0, param_epsilon=0.03, param_lambda=0, param_k=1,
        param_mu=0.001, param_pi=0, param_omicron=0.0, param_upsilon=1, device=0):
    """Run ChronoClust on a list of per-timepoint CSV files (in time order).  See the reference's
    app.run docstring for the meaning of every parameter; `device` (CUDA ordinal) is the only addition."""
    program_state_dir = '{}/program_images'.format(output_directory)
    logger = setup_logger('{}/logs'.format(output_directory))
    logger.info("Chronoclust start")

    config = {"beta": param_beta, "delta": param_delta, "epsilon": param_epsilon, "lambda": param_lambda,
              "k": param_k, "mu": param_mu, "pi": param_pi, "omicron": param_omicron, "upsilon": param_upsilon}

    program_state_dir_exists = os.path.exists(program_state_dir)
    if restore_program and program_state_dir_exists:
        logger.info("Restoring Chronoclust state saved in {}".format(program_state_dir))
        hddstream, tracker_by_association, tracker_by_lineage = restore_program_state(program_state_dir)
        hddstream.set_logger(logger)
        hddstream.set_config(config)
    else:
        if restore_program and not program_state_dir_exists:
            logger.warning("Restoring previous Chronoclust state not possible as program_images is not in {}".format(
                output_directory))
        logger.info("Setup new Chronoclust state")
        hddstream = HDDStream(config, logger, device=device)
        tracker_by_association = TrackByHistoricalAssociation()
        tracker_by_lineage = TrackByLineage()

    dataset_attributes = get_dataset_attributes(data[0])
    result_filename = f'{output_directory}/result.csv'
    result_file_header = ['timepoint', 'cumulative_size', 'pcore_ids', 'pref_dimensions'] + dataset_attributes + \
                         ['tracking_by_lineage', 'tracking_by_association']

    gating_df = None if gating_centroid_file is None else pd.read_csv(gating_centroid_file)
    gating = defaultdict(dict)
    if gating_df is not None:
        result_file_header.append('predicted_label')
        for _, gate in gating_df.iterrows():
            centroid = tuple(gate[dataset_attributes].values)
            gating[int(gate['Day'])][centroid] = gate['PopName']

    write_file_header(result_filename, result_file_header)

    scaler = None
    if normalise_data:
        logger.info("Setting up scaler")
        scaler = Scaler(data)

    for timepoint, data_file in enumerate(data):
        if restore_program and program_state_dir_exists and hddstream.last_data_timestamp >= timepoint:
            continue
        logger.info("Processing dataset {}".format(timepoint))
        dataset = pd.read_csv(data_file, header=0, sep=',').to_numpy()
        if normalise_data:
            logger.info("Scaling dataset {}".format(timepoint))
            dataset = scaler.scale_data(dataset)
        dataset = np.ascontiguousarray(dataset, dtype=np.float64)

        # ---- the hot path: online + offline clustering on the GPU
        hddstream.online_microcluster_maintenance(dataset, timepoint)
        pcore_by_id = {x.id[0]: x for x in hddstream.pcore_MC}

        for fc in hddstream.final_clusters:
            rounded_weight = Decimal(str(fc.cumulative_weight)).quantize(Decimal('1.1'), rounding=ROUND_HALF_UP)
            cluster = Cluster(list(fc.id), fc.cluster_centroids, rounded_weight, fc.preferred_dimension_vector)
            cluster.add_pcore_objects(pcore_by_id)
            tracker_by_lineage.add_new_child_cluster(cluster)

        tracker_by_lineage.calculate_ids()
        tracker_by_association.set_current_clusters(tracker_by_lineage.child_clusters)
        tracker_by_association.track_cluster_history()

        write_result_file(gating, result_filename, timepoint, tracker_by_association, scaler=scaler)
        write_datapoints_details(dataset_attributes, tracker_by_lineage.child_clusters, hddstream, dataset,
                                 f'{output_directory}/cluster_points_D{timepoint}.csv', scaler)

        tracker_by_lineage.transfer_child_to_parent()
        tracker_by_association.transfer_current_to_previous()

        logger.info("Saving Chronoclust state for timepoint {}".format(timepoint))
        save_program_state(hddstream, output_directory, tracker_by_association, tracker_by_lineage)

    with open(f'{output_directory}/parameters.csv', 'w') as f:
        w = csv.DictWriter(f, config.keys())
        w.writeheader()
        w.writerow(config)
    logger.info('Chronoclust finish')


def write_result_file(gating, result_filename, timepoint, tracker_by_association, scaler):
    """One row per cluster, same columns / formatting as the reference (app.py:229-260)."""
    result = []
    gating_now = gating.get(timepoint)
    for cluster in tracker_by_association.current_clusters:
        row = [timepoint, cluster.cumulative_weight, cluster.get_pcore_ids_as_str(),
               cluster.get_preferred_dimensions_as_str()]
        if scaler:
            centroid = scaler.reverse_scaling([cluster.centroid]).tolist()[0]
            centroid = np.round(centroid, 5).tolist()
        else:
            centroid = np.round(cluster.centroid, 5).tolist()
        row.extend(centroid)
        row.append(cluster.id)
        row.append(cluster.get_historical_associates_as_str())
        if bool(gating_now):
            row.append(find_closest_gating(gating_now, cluster, scaler))
        result.append(row)
    append_to_file(result_filename, result)


def write_datapoints_details(dataset_attributes, clusters, hddstream, dataset, cluster_points_filename, scaler):
    """cluster_points_D{t}.csv: id (input row), cluster_id (lineage label or None), marker values; rows in
    input order (app.py:263-360).  Built from the device's row -> MC assignment: row -> MC uid -> pcore id
    -> the cluster holding that pcore id."""
    n = dataset.shape[0]
    ids, uids, *_ = hddstream.export_arrays(0)
    label_of_pcore = {}
    for cluster in clusters:
        for pid in cluster.pcore_ids:
            label_of_pcore[pid] = cluster.id
    label_of_uid = {int(u): label_of_pcore.get(int(i), "None") for i, u in zip(ids, uids)}
    assign = hddstream.last_assignment
    uniq, inv = np.unique(assign, return_inverse=True)
    labels = np.array([label_of_uid.get(int(u), "None") for u in uniq], dtype=object)[inv]
    values = scaler.reverse_scaling(dataset) if scaler else dataset
    cols = {'id': np.arange(n), 'cluster_id': labels}
    for j, name in enumerate(dataset_attributes):
        cols[name] = values[:, j]
    pd.DataFrame(cols).to_csv(cluster_points_filename, index=False)


def save_program_state(hddstream, output_dir, tracker_by_association, tracker_by_lineage):
    d = "{}/program_images".format(output_dir)
    if not os.path.exists(d):
        os.mkdir(d)
    for name, obj in ((HDDSTREAM_OBJ, hddstream), (TRACKER_HISTORICAL_ASSOC, tracker_by_association),
                      (TRACKER_LINEAGE, tracker_by_lineage)):
        with open('{}/{}'.format(d, name), 'wb') as f:
            pickle.dump(obj, f)


def restore_program_state(program_state_dir):
    out = []
    for name in (HDDSTREAM_OBJ, TRACKER_HISTORICAL_ASSOC, TRACKER_LINEAGE):
        with open('{}/{}'.format(program_state_dir, name), 'rb') as f:
            out.append(pickle.load(f))
    return out[0], out[1], out[2]


def setup_logger(log_dir):
    if not os.path.exists(log_dir):
        os.makedirs(log_dir)
    logging.basicConfig(filename='{}/Chronoclust.log'.format(log_dir),
                        format='%(asctime)s [%(levelname)-8s] %(message)s')
    logger = logging.getLogger()
    logger.setLevel(logging.INFO)
    return logger


def write_file_header(filename, header):
    with open(filename, 'w') as f:
        csv.writer(f).writerow(header)


def append_to_file(filename, content):
    with open(filename, 'a') as f:
        csv.writer(f).writerows(content)


def get_dataset_attributes(dataset_file):
    return pd.read_csv(dataset_file, sep=',', header=None).iloc[0].values.tolist()


def find_closest_gating(gating_dict, cluster, scaler):
    best_d, best_label = None, None
    for centroid, label in gating_dict.items():
        centroid_norm = scaler.scale_data([centroid])[0].tolist() if scaler else centroid
        d = cluster.get_projected_dist_to_point(np.array(centroid_norm))
        if best_d is None or d < best_d:
            best_d, best_label = d, label
    return best_label
