"""Shared test helpers: rebuild the reference's golden frame with the oracle."""
import numpy as np

from oracle import ekf_oracle as O


def build_golden_frame(golden):
    """Recipe of SURVEY §4: initialize_x_and_p -> add the 13 features from their
    uv_when_initialized (rho0=1, std_rho=1, std_pxl=1) -> one prediction."""
    cam = O.initialize_cam()
    x, P = O.initialize_x_and_p()
    filt = O.ekf_filter(x, P, 0.007, 0.007, 1.0, "constant_velocity")
    features_info = []
    for k in range(golden["uv_when_initialized"].shape[0]):
        uv = golden["uv_when_initialized"][k]
        X, Pn, newf = O.add_features_inverse_depth(uv, filt.x_k_k, filt.p_k_k, cam, filt.std_z, 1.0, 1.0)
        filt.x_k_k, filt.p_k_k = X, Pn
        features_info.append(O.new_feature_info(uv, X, int(golden["init_frame"][k]), newf))
    filt, features_info = O.ekf_prediction(filt, features_info)
    return cam, filt, features_info
