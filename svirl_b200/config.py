"""Process-wide solver configuration (same role and names as svirl/config.py:1-36).

``GLSolver.__init__`` fills these module globals and every class reads them, so, as in the
reference, one solver per process is the supported use."""

Nx = dx = Lx = None
Ny = dy = Ly = None
Nz = dz = Lz = None
Nt = dt = T = None
NtA = dtA = TA = None

material_tiling = None
order_parameter = 'random'
random_seed = None
random_level = 1.0

normal_conductivity = 1.0
linear_coefficient = 1.0
gl_parameter = float('inf')

homogeneous_external_field = 0.0
external_field = 0.0

fixed_vortices = None
fixed_vortices_correction = 'cell centers'
phase_lock_radius = None

order_parameter_Langevin_coefficient = 0.0
vector_potential_Langevin_coefficient = 0.0

device_id = 0
dtype = None
dtype_complex = None
stop_criterion_order_parameter = 1e-6
stop_criterion_vector_potential = 1e-6
convergence_rtol = 1e-6
# new (not in the reference): 'reference' = SciPy BFGS on the raw coefficients with a rescue when it
# runs away, 'normalized' = always minimise c / max|c| (robust on very large grids)
cg_line_search = 'reference'

# slab decomposition (new): rows [j0, j1) of the global grid are owned by this process
slab = None
