"""Unit conversions -- host-side arithmetic identical to the reference (cardiax/convert.py:4-67)."""


def params_to_units(params, dx, dt):
    # convert.py:4-15
    params["Cm"] /= (dx ** 2) / dt
    for key in ("tau_d", "tau_v1_minus", "tau_v2_minus", "tau_v_plus", "tau_0", "tau_r", "tau_si", "tau_w_minus",
                "tau_w_plus"):
        params[key] /= dt
    return params


def diffusivity_to_units(d, dt):
    return d / dt


def diffusivity_rescale(c, domain):
    # convert.py:23-26
    a, b = c.min(), c.max()
    y, z = domain[0], domain[1]
    return (c - a) * (z - y) / (b - a) + y


def realsize_to_shape(field, dx):
    return (int(field[0] / dx), int(field[1] / dx))


def shape_to_realsize(field, dx):
    return (int(field[0] * dx), int(field[1] * dx))


def cm_to_units(value, dx):
    return int(value / dx)


def units_to_cm(value, dx):
    return value * dx


def ms_to_units(value, dt):
    return int(value / dt)


def units_to_ms(value, dt):
    return value * dt


def stimuli_to_units(stimuli, dx, dt):
    # convert.py:53-59
    stimuli = list(stimuli)
    for i in range(len(stimuli)):
        stimuli[i]["start"] = ms_to_units(stimuli[i]["start"], dt)
        stimuli[i]["duration"] = ms_to_units(stimuli[i]["duration"], dt)
        stimuli[i]["period"] = ms_to_units(stimuli[i]["period"], dt)
    return stimuli


def u_to_V(u, V0=-85, Vfi=15):
    return ((Vfi - V0) * u) + V0


def V_to_u(V, V0=-85, Vfi=15):
    return (V - V0) / (Vfi - V0)
