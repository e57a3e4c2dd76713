"""Named workloads as parameter overrides over examples/general.hydroconf."""
from hydro_b200.config import Params


def rt3d(n=8, fixed_work=True, **kw):
    """W4: Rayleigh-Taylor 3-D = general.hydroconf + examples/rt/mfer.hydroconf + MODULE hydro3d."""
    p = Params(
        MODULE="hydro3d", Nx=n, Ny=n, Nz=n, A=(0, 0, 0), B=(1, 1, 1), A1=(0, 0, 0), B1=(1, 0.5, 1),
        dt=0.005, dt_auto=0, T=20.0,
        initial_volume_fraction_smooth_times=0, density_smooth_times=0, viscosity_smooth_times=0,
        simpler=0, sharp=0.0, convergence_tolerance=1e-4, advection_dt_factor=1.0, lu_relaxed_tolerance=1e-5,
        num_iterations_limit=20, rhie_chow_factor=1.0, pressure_fixed_point=(0, 1, 0), pressure_fixed_value=0.0,
        guess_extrapolation=0.0, force=(0., 0, 0), gravity=(0., -1., 0), sigma=0.0,
        num_phases=2, density_0=2.0, density_1=1.0, viscosity_0=0.002, viscosity_1=0.001,
        cfl=0.25, cfl_advection=0.25, initial_velocity=(0, 0.01, 0), initial_sin_n=(1, 0, 0),
        initial_sin_lambda=0.4, initial_sin_phase=1.57079632679,
    )
    if fixed_work:
        p.update(num_iterations_limit=3, convergence_tolerance=0.0, lu_relaxed_num_iters_limit=100,
                 lu_relaxed_tolerance=0.0)
    p.update(kw)
    return p


def cavity(n=32, **kw):
    """W1: examples/cavity (2-D lid-driven cavity, Re 3200)."""
    p = Params(Nx=n, Ny=n, T=1.0, dt=0.01, viscosity_0=0.0003125, condition_top="wall 1 0 0",
               velocity_relaxation_factor=0.8, pressure_relaxation_factor=0.9, num_iterations_limit=1,
               convergence_tolerance=1e-6, lu_relaxed_num_iters_limit=1000, lu_relaxed_tolerance=0.0)
    p.update(kw)
    return p


def cavity_kat():
    """The reference's golden sample: examples/cavity/sample/exp.log:5-170."""
    return cavity(64, T=1e5, dt=1e5, num_iterations_limit=1000, convergence_tolerance=1e-6,
                  lu_relaxed_num_iters_limit=100, lu_relaxed_tolerance=0.0)


def broken_dam_2d(nx=142, ny=80, **kw):
    """W2: examples/broken_dam_2d."""
    p = Params(A=(0, 0, 0), B=(1.25, 0.7, 1), A1=(0, 0, 0), B1=(0.2, 0.25, 1), Nx=nx, Ny=ny, T=0.6, dt=0.0025,
               num_phases=2, gravity=(0., -10, 0), density_0=1.255, density_1=1000., viscosity_0=1.7e-05,
               viscosity_1=0.001, initial_volume_fraction_smooth_times=2, density_smooth_times=3,
               viscosity_smooth_times=3)
    p.update(kw)
    return p


def broken_dam_3d(nx=64, ny=20, nz=20, **kw):
    """W5: examples/broken_dam_3d (obstacle box -> excluded cells)."""
    p = Params(MODULE="hydro3d", A=(0, 0, 0), B=(3.2, 1, 1), A1=(0, 0, 0), B1=(1.2, 1, 0.55), Nx=nx, Ny=ny, Nz=nz,
               T=8.0, dt=0.0025, num_phases=2, gravity=(0., 0, -10), density_0=1.255, density_1=1000.,
               viscosity_0=1.7e-05, viscosity_1=0.001, box_A=(2.37, 0.3, 0), box_B=(2.53, 0.7, 0.16),
               initial_volume_fraction_smooth_times=2, density_smooth_times=3, viscosity_smooth_times=3,
               num_iterations_limit=5)
    p.update(kw)
    return p


def thermal_2d(nx=32, ny=16, **kw):
    """W3: two-phase drop in a channel with the temperature equation switched on
    (examples/mortazavi parameters + heat_enable 1)."""
    p = Params(A=(0, 0, 0), B=(2.0, 1.0, 1.0), Nx=nx, Ny=ny, dt=0.01, num_phases=2,
               IC=(0.5, 0.4, 0.0), IR=0.2, density_0=1.0, density_1=2.0, viscosity_0=0.05, viscosity_1=0.1,
               conductivity_0=0.01, conductivity_1=0.05, force=(1.0, 0, 0), sigma=0.0,
               initial_volume_fraction_smooth_times=1, density_smooth_times=1, viscosity_smooth_times=1,
               heat_enable=1, heat_box_lb=(-1, -1, -1), heat_box_rt=(3, 0.01, 1), heat_box_temperature=1.0,
               temperature_initial=0.0, num_iterations_limit=4, convergence_tolerance=1e-5,
               lu_relaxed_num_iters_limit=50, lu_relaxed_tolerance=1e-6, advection_dt_factor=0.5,
               condition_left="wall 0 0 0", condition_top="wall 0.5 0 0")
    p.update(kw)
    return p


# Cases whose result is a badly conditioned function of the state: meshvel_auto vcx differences the centre of phase 1 over one
# step (hydro2d.hpp:1459-1463), which amplifies the last-bit differences of the centre (summation order of the statistics,
# cell centres as node averages in the reference) by cx / (cx - previous cx) ~ 1e4 -- into the mesh velocity and from there
# into every flux.  name -> tolerance of the field comparisons
ILL_CONDITIONED = {"dam3d_24x8x8_meshvel_vcx": 1e-9}

# fixtures tests/golden/ref_<name>.npz are dumped from the real reference by oracle/make_golden.py
# name -> (Params, nsteps)
GOLDEN_CASES = {
    "rt3d_8": (rt3d(8), 2),
    "rt3d_16": (rt3d(16), 2),
    "rt3d_12x10x9": (rt3d(8, Nx=12, Ny=10, Nz=9, B=(1.3, 0.9, 1.1)), 2),
    "rt3d_8_asconfigured": (rt3d(8, fixed_work=False, lu_relaxed_num_iters_limit=200), 2),
    "rt3d_8_jacobi_dtauto": (rt3d(8, dt_auto=1, linear_solver_pressure="jacobi", lu_relaxed_relaxation_factor=0.9,
                                        lu_relaxed_tolerance=1e-4, convergence_tolerance=1e-4,
                                        num_iterations_limit=6), 3),
    "cavity_16": (cavity(16, num_iterations_limit=5, lu_relaxed_num_iters_limit=30), 3),
    "cavity_12_lurelaxed": (cavity(12, num_iterations_limit=4, linear_solver_pressure="lu_relaxed",
                                         lu_relaxed_num_iters_limit=10, lu_relaxed_relaxation_factor=0.5,
                                         time_second_order=0, guess_extrapolation=0.5, meshvel=(0.1, 0, 0)), 3),
    "dam2d_36x20": (broken_dam_2d(36, 20, lu_relaxed_num_iters_limit=50), 3),
    "dam3d_32x10x10": (broken_dam_3d(32, 10, 10, lu_relaxed_num_iters_limit=40), 2),
    "thermal2d_32x16": (thermal_2d(), 3),
    # the linear-solver factory (hydro2d.hpp:194-218) applies to every system: velocity, pressure, temperature
    "cavity_12_velgs_plu": (cavity(12, num_iterations_limit=4, linear_solver_velocity="gauss_seidel", linear_solver_pressure="lu",
                                   lu_relaxed_num_iters_limit=25, lu_relaxed_tolerance=1e-9), 3),
    "rt3d_8_veljacobi": (rt3d(8, linear_solver_velocity="jacobi", lu_relaxed_num_iters_limit=40, lu_relaxed_tolerance=1e-7,
                              lu_relaxed_relaxation_factor=0.9), 2),
    # interface sharpening (advection.hpp:479-529), with and without the directional split
    "rt3d_8_sharp_split": (rt3d(8, sharp=0.05, tvd_split=1), 3),
    "dam2d_32x16_sharp": (broken_dam_2d(32, 16, sharp=0.02, lu_relaxed_num_iters_limit=40), 3),
    # outlet conditions with the outlet mass balance (fluid.hpp:309-336, 542-600)
    "channel2d_32x16_outlet": (cavity(16, Nx=32, Ny=16, B=(2, 1, 1), condition_top="wall 0 0 0", condition_left="inlet 1 0 0",
                                      condition_right="outlet", num_iterations_limit=5, lu_relaxed_num_iters_limit=60,
                                      viscosity_0=0.01), 3),
    "rt3d_8_inlet_outlet": (rt3d(8, condition_left="inlet 0.05 0 0", condition_right="outlet"), 2),
    # phase slip (Stokes settling relative to the mixture, hydro2d.hpp:1030-1122)
    "rt3d_8_settling": (rt3d(8, enable_settling_1=1, bubble_radius_1=0.02, initial_volume_fraction_smooth_times=2), 3),
    "dam2d_32x16_settling": (broken_dam_2d(32, 16, enable_settling_0=1, bubble_radius_0=0.0004, enable_settling_1=1,
                                           bubble_radius_1=0.0002, lu_relaxed_num_iters_limit=40), 3),
    # SIMPLER: second pressure solve per iteration (fluid.hpp:1060-1155)
    "rt3d_8_simpler": (rt3d(8, simpler=1), 2),
    "cavity_16_simpler": (cavity(16, simpler=1, num_iterations_limit=5, lu_relaxed_num_iters_limit=30), 3),
    "thermal2d_24x12_vellur_heatgs": (thermal_2d(24, 12, linear_solver_velocity="lu_relaxed", linear_solver_heat="gauss_seidel",
                                                 lu_relaxed_num_iters_limit=12, lu_relaxed_relaxation_factor=0.7), 2),
    # automatic mesh velocity: the mesh follows phase 1 (CalcStat, hydro2d.hpp:1510-1524; examples/mortazavi, coal3d, epfl)
    "dam2d_32x16_meshvel_vx": (broken_dam_2d(32, 16, meshvel_auto="vx", meshvel=(0.05, 0, 0), lu_relaxed_num_iters_limit=40), 4),
    "dam3d_24x8x8_meshvel_vcx": (broken_dam_3d(24, 8, 8, meshvel_auto="vcx", meshvel_weight=0.3, lu_relaxed_num_iters_limit=30), 4),
}


