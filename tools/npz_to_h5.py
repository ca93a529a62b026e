"""Convert a per-chain result file written by rfsurfhmc_b200 (.npz) into the reference's HDF5 tree
(/root/reference/pyhmc/hmc.py:203-226,272-275) so that the reference's plot.py can read it.
Needs h5py (not available in the build image):

    python tools/npz_to_h5.py results/chain_joint.0.npz   ->   results/chain_joint.0.h5
"""
import sys
import numpy as np


def convert(path):
    import h5py
    z = np.load(path)
    out = path[:-4] + ".h5"
    with h5py.File(out, "w") as f:
        f.create_dataset("initmodel", data=z["initmodel"])
        f.create_dataset("obs", data=z["obs"])
        f.create_dataset("mean/model", data=z["mean/model"])
        f.create_dataset("mean/syn", data=z["mean/syn"])
        for i in range(z["models"].shape[0]):
            f.create_dataset(f"{i}/model", data=z["models"][i])
            f.create_dataset(f"{i}/syn", data=z["syn"][i])
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(convert(p))
