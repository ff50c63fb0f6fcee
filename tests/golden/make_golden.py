"""Extracts the numeric fields of the reference's only golden vector,
/root/reference/matlab_code/features_information.mat (a features_info snapshot taken
after search_IC_matches on the first processed frame: 13 inverse-depth features, n=91),
into tests/golden/features_information.npz.  Image patches are dropped.

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
"""
import os
import numpy as np
import scipy.io as sio

SRC = "/root/reference/matlab_code/features_information.mat"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "features_information.npz")


def main():
    m = sio.loadmat(SRC, squeeze_me=False, struct_as_record=False)
    fi = m["features_info"]
    N = fi.shape[1]
    out = dict(
        uv_when_initialized=np.zeros((N, 2)), yi=np.zeros((N, 6)), z=np.zeros((N, 2)),
        h=np.zeros((N, 2)), S=np.zeros((N, 2, 2)), R=np.zeros((N, 2, 2)),
        individually_compatible=np.zeros(N, dtype=np.int32),
        low_innovation_inlier=np.zeros(N, dtype=np.int32),
        high_innovation_inlier=np.zeros(N, dtype=np.int32),
        init_frame=np.zeros(N, dtype=np.int32), times_predicted=np.zeros(N, dtype=np.int32),
    )
    Hs = []
    types = []
    for i in range(N):
        f = fi[0, i]
        out["uv_when_initialized"][i] = np.asarray(f.uv_when_initialized, dtype=np.float64).reshape(2)
        out["yi"][i] = np.asarray(f.yi, dtype=np.float64).reshape(6)
        out["z"][i] = np.asarray(f.z, dtype=np.float64).reshape(2)
        out["h"][i] = np.asarray(f.h, dtype=np.float64).reshape(2)
        out["S"][i] = np.asarray(f.S, dtype=np.float64)
        out["R"][i] = np.asarray(f.R, dtype=np.float64)
        for k in ("individually_compatible", "low_innovation_inlier", "high_innovation_inlier",
                  "init_frame", "times_predicted"):
            out[k][i] = int(np.asarray(getattr(f, k)).reshape(-1)[0])
        Hs.append(np.asarray(f.H.todense(), dtype=np.float64))
        types.append(str(np.asarray(f.type).reshape(-1)[0]))
    out["H"] = np.stack(Hs)
    out["type"] = np.array(types)
    np.savez_compressed(DST, **out)
    print("wrote", DST, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
