"""Helpers shared by the golden-vector tests: load a fixture written by oracle/gen_golden.py."""
import glob
import json
import os

import numpy as np

from fleetrl_b200._abi import FleetConsts

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.consts_dict = json.loads(str(z["consts_json"]))
        self.meta = json.loads(str(z["meta_json"]))
        self.tables = {k[3:]: z[k] for k in z.files if k.startswith("tb_")}
        self.traj = {k[3:]: z[k] for k in z.files if k.startswith("tr_")}
        self.log = {k[4:]: z[k] for k in z.files if k.startswith("log_")}     # DataLogger columns (log_data=True cases)
        self.actions = z["actions"]
        self.start_idx = z["start_idx"]
        self.ep_start_rows = z["ep_start_rows"]
        self.n_steps_per_ep = int(z["n_steps_per_ep"])

    def consts(self, **over) -> FleetConsts:
        d = dict(self.consts_dict)
        d.update(over)
        return FleetConsts.from_dict(d)
