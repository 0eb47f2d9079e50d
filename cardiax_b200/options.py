"""Run-time knobs of the B200 solver (module-level, read at every call).

numerics           "fast" (default): reciprocal multiplications + FMAs, a few ulp per step from the
                   reference's operation order; "exact": the reference's order, op for op, bit-identical
                   to the CPU oracle and to the reference's own source (tests/test_reference_pin.py); slower -- every
                   division keeps the reference's rounding (see DESIGN.md section 6 for the measured ratio).
steps_per_launch   temporal blocking depth T (0 = library default).
kernel             0 auto, 1 general tile kernel only, 2 require the streaming kernel, 3 require the one-step wide kernel,
                   4 require the resident kernel (whole call in one cooperative launch, state in shared memory),
                   5 require its cluster form (a tissue of <= 16 tiles = one thread-block cluster, halos through DSMEM).
tiles              resident kernel: (rows, columns) of the tile grid, one CTA per tile ((0, 0) = planner's choice).
"""
numerics = "fast"
steps_per_launch = 0
kernel = 0
cta_threads = 0
rows_per_cta = 0
tiles = (0, 0)
cells_per_thread = 0   # resident kernel: 1, 2 or 4 (0 = planner's choice)
maps_global = 0        # resident kernel: 1 = diffusivity maps read from L2 instead of shared memory (0 = only if needed)
edge_tile = (0, 0)     # resident kernel: (rows, 4-column groups) of the tiles at the tissue's edges (0 auto, < 0 even)
detect_uniform_diffusivity = True
safe_division = False   # exact numerics: force every division through __fdiv_rn
verbose = True
# TimeIntegrator.DORMANDPRINCE: the keyword defaults of jax.experimental.ode.odeint, which the reference never overrides
ode_rtol = 1.4e-8
ode_atol = 1.4e-8
ode_mxstep = float("inf")
