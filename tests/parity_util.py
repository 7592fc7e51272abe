"""Shared helpers for the GPU-vs-oracle parity tests."""
import numpy as np

# Scaled error: |a - b| / max(|b|, floor). The floor keeps components that sit at zero in trimmed flight
# (angular rates, lateral velocity, friction-regulator states) from turning rounding noise into an
# "infinite" relative error; floors are small against each block's natural magnitude.
def state_floors(nx, variant, kin):
    nact = 0 if variant == 0 else (14 if variant == 4 else 7)   # variant 4: Actuator2 units, (v, p) per unit
    f = np.empty(nx)
    f[0:2] = 1e-2          # alpha_filt, beta_filt [rad]
    f[2:8] = 1e-2          # contact friction regulators [-]
    f[8] = 1e-2            # fuel fraction
    f[9] = 1.0             # engine omega [rad/s]
    f[10:12] = 1e-2        # idle / friction compensators
    k0 = 12 + nact
    f[12:k0] = 1e-2        # actuator positions
    if variant == 4:
        f[12:k0:2] = 1.0   # actuator velocities [1/s]: a 0.1 command step peaks near 2
    nkin = {0: 9, 1: 8, 2: 6}[kin]
    if kin == 0:
        f[k0:k0 + 8] = 1e-2   # unit quaternions
        f[k0 + 8] = 1.0       # h_e [m]
    elif kin == 1:
        f[k0:k0 + 7] = 1e-2   # unit quaternion q_eb, n-vector
        f[k0 + 7] = 1.0       # h_e [m]
    elif kin == 2:
        f[k0:k0 + 5] = 1e-2   # Euler angles, lat, lon [rad]
        f[k0 + 5] = 1.0
    d0 = k0 + nkin
    f[d0:d0 + 3] = 1e-2    # omega_eb_b [rad/s]
    f[d0 + 3:d0 + 6] = 1.0  # v_eb_b [m/s]
    return f


def scaled_err(a, b, floors):
    return np.abs(a - b) / np.maximum(np.abs(b), floors[:, None] if b.ndim == 2 else floors)
