"""NumPy restatement of what surrounds the Euler loop in the reference.  TEST INFRASTRUCTURE ONLY.

* ``odeint_dopri5`` -- ``cardiax/solve.py:88-89, 114-124``: ``jax.experimental.ode.odeint(step, state, ts, params,
  diffusivity, stimuli, dx)``.  jax is an UN-VENDORED dependency of the reference (``install_jax.sh:2`` pins jaxlib
  0.1.64; jax itself is unpinned, spring 2021, i.e. 0.2.1x).  The published algorithm of ``jax/experimental/ode.py`` of
  that vintage is restated here function by function (``initial_step_size``, ``runge_kutta_step``, ``error_ratio`` =
  mean of squares, ``optimal_step_size``, ``interp_fit_dopri`` / ``fit_4th_order_polynomial``, ``_odeint.scan_fun``).
  fp32 throughout (jax's default): array intermediates rounded per operation, the scalar controller in fp32 too.  What
  cannot be restated: XLA's summation order in ``jnp.dot`` / ``jnp.mean`` / ``jnp.linalg.norm`` and its fp32 ``pow``;
  here ``dot`` accumulates left to right, reductions accumulate in fp64 and round once, ``pow`` is the fp64 libm one
  rounded once.  PARITY UNPINNED (no jax here, no golden vector in the reference for this path).
* ``resize_bilinear`` -- ``cardiax/io.py:118-124``: ``jax.image.resize(a, shape, "bilinear")`` with its default
  anti-aliasing (``jax/_src/image/scale.py``: ``compute_weight_mat`` + one ``einsum``), evaluated in fp64.
* ``electrogram`` -- ``cardiax/metrics.py:13-22``.
"""
import numpy as np

from .fk_oracle import State, step

F = np.float32


# --------------------------------------------------------------------------- jax.experimental.ode (Dopri5)
_ALPHA = [F(x) for x in (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1., 0)]
_BETA = [[F(x) for x in row] for row in (
    [1 / 5, 0, 0, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0, 0],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0])]
_C_SOL = [F(x) for x in (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0)]
_C_ERR = [F(x) for x in (35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
                         -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.)]
_C_MID = [F(x) for x in (6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
                         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)]


def _dot(c, k):
    """jnp.dot(c, k) over the stages present in k: fp32 products accumulated left to right (zero weights add nothing)."""
    acc = c[0] * k[0]
    for j in range(1, len(k)):
        if c[j] != 0:
            acc = acc + c[j] * k[j]
    return acc


def _norm(x):
    """jnp.linalg.norm: sqrt(sum |x|^2) -- squares in fp32, sum in fp64 rounded once, fp32 sqrt."""
    return np.sqrt(F(np.sum((x * x).astype(np.float64))))


def _pow(a, b):
    return F(np.power(np.float64(a), np.float64(F(b))))


def _ravel(state):
    return np.concatenate([np.asarray(x, F).reshape(-1) for x in state])


def _optimal_step_size(last_step, mean_error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5.0):
    dfactor = F(1.0) if mean_error_ratio < 1 else F(dfactor)
    err_ratio = np.sqrt(mean_error_ratio)
    factor = max(F(1.0 / ifactor), min(_pow(err_ratio, 1.0 / order) / F(safety), F(1.0) / dfactor))
    return last_step * F(ifactor) if mean_error_ratio == 0 else last_step / factor


def odeint_dopri5(state, ts, params, diffusivity, stimuli, dx, rtol=1.4e-8, atol=1.4e-8, mxstep=np.inf, tanh="xla",
                  stats=None):
    """-> State of stacked arrays (len(ts), H, W): odeint(step, state, ts, params, diffusivity, stimuli, dx)."""
    shape = np.asarray(state[0]).shape
    n = int(np.prod(shape))
    rtol, atol = F(rtol), F(atol)
    ts = np.asarray(ts, F)
    evals = [0]

    def func(y, t):
        evals[0] += 1
        st = State(*[y[i * n:(i + 1) * n].reshape(shape) for i in range(3)])
        return _ravel(step(st, F(t), params, diffusivity, stimuli, dx, tanh=tanh))

    y0 = _ravel(state)
    with np.errstate(all="ignore"):
        f0 = func(y0, ts[0])
        # initial_step_size(fun, t0, y0, order=4, rtol, atol, f0)
        scale = atol + np.abs(y0) * rtol
        d0, d1 = _norm(y0 / scale), _norm(f0 / scale)
        h0 = F(1e-6) if (d0 < F(1e-5) or d1 < F(1e-5)) else F(0.01) * d0 / d1
        y1 = y0 + h0 * f0
        f1 = func(y1, ts[0] + h0)
        d2 = _norm((f1 - f0) / scale) / h0
        if d1 <= F(1e-15) and d2 <= F(1e-15):
            h1 = max(F(1e-6), h0 * F(1e-3))
        else:
            h1 = _pow(F(0.01) / (d1 + d2), 1. / (4 + 1.))
        dt = min(F(100.) * h0, h1)

        y, f, t, last_t = y0, f0, ts[0], ts[0]
        coeff = [y0] * 5
        out = [y0]
        attempts = accepted = 0
        for target in ts[1:]:
            i = 0
            while t < target and i < mxstep and dt > 0:
                # runge_kutta_step
                k = [f]
                for s in range(1, 7):
                    ti = t + dt * _ALPHA[s - 1]
                    yi = y + dt * _dot(_BETA[s - 1], k)
                    k.append(func(yi, ti))
                ny = dt * _dot(_C_SOL, k) + y
                err = dt * _dot(_C_ERR, k)
                nt = t + dt
                # error_ratio: mean((err / (atol + rtol * max(|y0|, |y1|))) ** 2)
                tol = atol + rtol * np.maximum(np.abs(y), np.abs(ny))
                q = err / tol
                ratio = F(np.sum((q * q).astype(np.float64))) / F(3 * n)
                # interp_fit_dopri
                y_mid = y + dt * _dot(_C_MID, k)
                dy0, dy1 = k[0], k[6]
                a = F(-2.) * dt * dy0 + F(2.) * dt * dy1 - F(8.) * y - F(8.) * ny + F(16.) * y_mid
                b = F(5.) * dt * dy0 - F(3.) * dt * dy1 + F(18.) * y + F(14.) * ny - F(32.) * y_mid
                c = F(-4.) * dt * dy0 + dt * dy1 - F(11.) * y - F(5.) * ny + F(16.) * y_mid
                new_coeff = [a, b, c, dt * dy0, y]
                new_dt = _optimal_step_size(dt, ratio)
                attempts += 1
                if ratio <= 1.:
                    accepted += 1
                    y, f, last_t, t, coeff = ny, k[6], t, nt, new_coeff
                dt = new_dt
                i += 1
            r = (target - last_t) / (t - last_t)
            yt = coeff[0]
            for cj in coeff[1:]:
                yt = yt * r + cj
            out.append(yt.astype(F))
    if stats is not None:
        stats.update(attempts=attempts, accepted=accepted, rhs_evals=evals[0])
    out = np.stack(out)
    return State(*[out[:, i * n:(i + 1) * n].reshape((len(ts),) + shape) for i in range(3)])


# --------------------------------------------------------------------------- jax.image.resize(..., "bilinear")
def resize_weights(n_in, n_out):
    """(n_in, n_out) fp64 weight matrix of compute_weight_mat (triangle kernel, antialias=True)."""
    inv_scale = n_in / n_out
    kernel_scale = max(inv_scale, 1.0)
    sample_f = (np.arange(n_out, dtype=np.float64) + 0.5) * inv_scale - 0.5
    x = np.abs(sample_f[None, :] - np.arange(n_in, dtype=np.float64)[:, None]) / kernel_scale
    w = np.maximum(0.0, 1.0 - x)
    total = w.sum(axis=0, keepdims=True)
    w = np.where(np.abs(total) > 1000.0 * np.finfo(np.float32).eps, w / np.where(total != 0, total, 1), 0.0)
    inside = (sample_f >= -0.5) & (sample_f <= n_in - 0.5)
    return np.where(inside[None, :], w, 0.0)


def resize_bilinear(a, size):
    """cardiax/io.py:118-124 -- resize the last two axes of ``a`` to ``size`` (fp64 evaluation)."""
    a = np.asarray(a, np.float64)
    wh, ww = resize_weights(a.shape[-2], size[0]), resize_weights(a.shape[-1], size[1])
    return np.swapaxes(np.swapaxes(a @ ww, -1, -2) @ wh, -1, -2)


# --------------------------------------------------------------------------- metrics.electrogram
def electrogram(x, point):
    """cardiax/metrics.py:13-22 (the reference's ogrid is [:W, :H]: square frames only)."""
    x = np.asarray(x, F)
    c_y, c_x = np.ogrid[: x.shape[-1], : x.shape[-2]]
    dist = np.sqrt(((c_x - point[0]) ** 2 + (c_y - point[1]) ** 2).astype(F))
    return F(1) * np.sum((x * dist).astype(np.float64), axis=(-1, -2)).astype(F)
