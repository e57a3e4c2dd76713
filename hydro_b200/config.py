"""Parameter handling for the GPU path: the reference's parameter names are the API.

`Params` accepts the same keys as the reference's per-experiment stores
P_int / P_double / P_bool / P_string / P_vect (control/experiment.hpp:50-54) and the
same `set <type> <name> <value>` lines as a .hydroconf script
(control/console.cpp:276-314).  Defaults are those of examples/general.hydroconf.
`Params.to_struct()` produces the `hg_config` of include/hydro_gpu.h.
"""
import ctypes as C
import re

HG_MAX_PHASES = 3
LINEAR_SOLVERS = {"lu": 0, "lu_relaxed": 1, "gauss_seidel": 2, "jacobi": 3}
BC_KINDS = {"wall": 0, "inlet": 1, "outlet": 2}
MESHVEL_AUTO = {"vx": 1, "vcx": 2}
SIDES = ["left", "right", "bottom", "top", "close", "far"]

d3 = C.c_double * 3
dP = C.c_double * HG_MAX_PHASES


class HgConfig(C.Structure):
    """ctypes mirror of `struct hg_config` (include/hydro_gpu.h); keep in sync."""
    _fields_ = [
        ("dim", C.c_int), ("Nx", C.c_int), ("Ny", C.c_int), ("Nz", C.c_int),
        ("A", d3), ("B", d3), ("box_A", d3), ("box_B", d3),
        ("condition_kind", C.c_int * 6), ("condition_velocity", d3 * 6),
        ("pressure_fixed_enable", C.c_int), ("pressure_fixed_point", d3),
        ("pressure_fixed_value", C.c_double),
        ("initial_velocity", d3), ("initial_pois", C.c_int), ("initial_sin_enable", C.c_int),
        ("initial_sin_n", d3), ("initial_sin_lambda", C.c_double), ("initial_sin_phase", C.c_double),
        ("A1", d3), ("B1", d3), ("A2", d3), ("B2", d3),
        ("IC", d3), ("IR", C.c_double), ("IC2", d3), ("IR2", C.c_double),
        ("initial_volume_fraction", dP), ("initial_volume_fraction_smooth_times", C.c_int),
        ("dt", C.c_double), ("dt_auto", C.c_int), ("cfl", C.c_double), ("cfl_advection", C.c_double),
        ("num_phases", C.c_int), ("density", dP), ("viscosity", dP), ("conductivity", dP),
        ("gravity", d3), ("force", d3), ("sigma", C.c_double),
        ("fluid_enable", C.c_int), ("advection_enable", C.c_int),
        ("convergence_tolerance", C.c_double), ("num_iterations_limit", C.c_int),
        ("velocity_relaxation_factor", C.c_double), ("pressure_relaxation_factor", C.c_double),
        ("rhie_chow_factor", C.c_double),
        ("time_second_order", C.c_int), ("simpler", C.c_int), ("force_geometric_average", C.c_int),
        ("guess_extrapolation", C.c_double), ("meshvel", d3), ("meshvel_output", C.c_int),
        ("linear_solver_velocity", C.c_int), ("linear_solver_pressure", C.c_int),
        ("linear_solver_heat", C.c_int),
        ("lu_relaxed_tolerance", C.c_double), ("lu_relaxed_num_iters_limit", C.c_int),
        ("lu_relaxed_relaxation_factor", C.c_double),
        ("density_smooth_times", C.c_int), ("viscosity_smooth_times", C.c_int),
        ("force_smooth_times", C.c_int),
        ("advection_dt_factor", C.c_double), ("tvd_split", C.c_int), ("sharp", C.c_double),
        ("heat_enable", C.c_int), ("temperature_initial", C.c_double),
        ("heat_box_lb", d3), ("heat_box_rt", d3), ("heat_box_temperature", C.c_double),
        ("heat_relaxation_factor", C.c_double), ("time_second_order_heat", C.c_int),
        ("world_size", C.c_int), ("rank", C.c_int), ("device", C.c_int),
        ("pressure_sweeps_per_check", C.c_int), ("solver_ctas", C.c_int),
        ("enable_settling", C.c_int * HG_MAX_PHASES), ("velocity_is_carrier", C.c_int), ("bubble_radius", dP),
        ("meshvel_auto", C.c_int), ("reserved0", C.c_int), ("meshvel_weight", C.c_double),
        ("reserved", C.c_int * 2),
    ]


class HgStepStats(C.Structure):
    """ctypes mirror of `struct hg_step_stats`."""
    _fields_ = [
        ("simple_iterations", C.c_int), ("convergence_indicator", C.c_double),
        ("pressure_sweeps_total", C.c_int), ("advection_substeps", C.c_int),
        ("dt", C.c_double), ("time", C.c_double), ("pressure_last_diff", C.c_double),
        ("volume", dP), ("mass", dP), ("pd_min", dP), ("pd_max", dP),
        ("center", d3 * HG_MAX_PHASES), ("velocity", d3 * HG_MAX_PHASES),
    ]


# field ids (enum hg_field)
F = dict(
    VELOCITY_X=0, VELOCITY_Y=1, VELOCITY_Z=2, PRESSURE=3, VOLUME_FLUX=4,
    PARTIAL_DENSITY_0=5, PARTIAL_DENSITY_1=6, PARTIAL_DENSITY_2=7, TEMPERATURE=8,
    DENSITY=9, VISCOSITY=10, FORCE_X=11, FORCE_Y=12, FORCE_Z=13,
    VOLUME_FRACTION_0=14, VOLUME_FRACTION_1=15, VOLUME_FRACTION_2=16,
    STFORCE_X=17, STFORCE_Y=18, STFORCE_Z=19,
    VELOCITY_PREV_X=20, VELOCITY_PREV_Y=21, VELOCITY_PREV_Z=22, PRESSURE_PREV=23,
    VOLUME_FLUX_PREV=24, EXCLUDED=25, CONDUCTIVITY=26,
)
FACE_FIELDS = {F["VOLUME_FLUX"], F["VOLUME_FLUX_PREV"]}

# examples/general.hydroconf: every key with its store type (the reference's module ctor
# throws "'k' undefined" for any missing one, common/data_structures.hpp:238-242)
GENERAL_DEFAULTS = {
    'MODULE': 'hydro2D_uniform_MPI',
    'plt_title': 'Template',
    'A': (0.0, 0.0, 0.0),
    'B': (1.0, 1.0, 1.0),
    'A1': (0.0, 0.0, 0.0),
    'B1': (0.0, 0.0, 1.0),
    'A2': (0.0, 0.0, 0.0),
    'B2': (0.0, 0.0, 1.0),
    'box_A': (0.0, 0.0),
    'box_B': (0.0, 0.0),
    'IC': (0.0, 0.0, 0.0),
    'IR': 0.0,
    'IC2': (0.0, 0.0, 0.0),
    'IR2': 0.0,
    'Nx': 100,
    'Ny': 100,
    'Nz': 5,
    'T': 1.0,
    'dt': 0.01,
    'dt_auto': 0,
    'cfl': 0.5,
    'cfl_advection': 0.5,
    'num_phases': 1,
    'gravity': (0.0, 0.0, 0.0),
    'force': (0.0, 0.0, 0.0),
    'sigma': 0.0,
    'density_0': 1.0,
    'density_1': 1.0,
    'density_2': 1.0,
    'viscosity_0': 1.0,
    'viscosity_1': 1.0,
    'viscosity_2': 1.0,
    'molar_0': 1.0,
    'molar_1': 1.0,
    'molar_2': 1.0,
    'deforming_velocity': 0,
    'initial_velocity': (0.0, 0.0),
    'condition_top': 'wall 0 0 0',
    'condition_bottom': 'wall 0 0 0',
    'condition_left': 'wall 0 0 0',
    'condition_right': 'wall 0 0 0',
    'condition_close': 'wall 0 0 0',
    'condition_far': 'wall 0 0 0',
    'chemistry': 'steady',
    'chem_intensity': 0.0,
    'reaction_zone_lb': (0.0, 0.0, 0.0),
    'reaction_zone_rt': (0.0, 0.0, 0.0),
    'radiation_enable': 0,
    'radiation_intensity': 1.0,
    'radiation_direction': (0.0, -1.0, 0.0),
    'radiation_box_lb': (0.07, 0.04, -0.01),
    'radiation_box_rt': (0.13, 0.06, 0.01),
    'absorption_rate_0': 100.0,
    'absorption_rate_1': 0.0,
    'absorption_rate_2': 0.0,
    'heat_enable': 0,
    'linear_solver_heat': 'lu',
    'heat_box_lb': (0.0, 0.0, 0.0),
    'heat_box_rt': (0.0, 0.0, 0.0),
    'conductivity_0': 1.0,
    'conductivity_1': 1.0,
    'conductivity_2': 1.0,
    'temperature_initial': 0.0,
    'heat_box_temperature': 0.0,
    'heat_relaxation_factor': 1.0,
    'time_second_order_heat': 1,
    'incompressible_relaxation': 0.0,
    'temperature_expansion_rate_0': 0.0,
    'temperature_expansion_rate_1': 0.0,
    'temperature_expansion_rate_2': 0.0,
    'temperature_expansion_base_0': 0.0,
    'temperature_expansion_base_1': 0.0,
    'temperature_expansion_base_2': 0.0,
    'fluid_enable': 1,
    'advection_enable': 1,
    'advection_solver': 'tvd',
    'advection_dt_factor': 0.1,
    'tvd_split': 0,
    'convergence_tolerance': 0.01,
    'num_iterations_limit': 10,
    'velocity_relaxation_factor': 0.8,
    'pressure_relaxation_factor': 0.9,
    'linear_solver_velocity': 'lu',
    'linear_solver_pressure': 'gauss_seidel',
    'lu_relaxed_relaxation_factor': 1.9,
    'lu_relaxed_num_iters_limit': 1000,
    'lu_relaxed_tolerance': 0.001,
    'time_second_order': 1,
    'rhie_chow_factor': 1.0,
    'simpler': 0,
    'initial_volume_fraction_smooth_times': 2,
    'density_smooth_times': 2,
    'viscosity_smooth_times': 2,
    'force_smooth_times': 0,
    'force_geometric_average': 0,
    'guess_extrapolation': 0.0,
    'compressible_enable': 0,
    'meshvel': (0.0, 0.0, 0.0),
    'meshvel_output': 1,
    'meshvel_weight': 0.5,
    'sharp': 0.0,
    'spawning_gap': 0.25,
    'particle_radius': 1.5,
    'min_num_particles': 3,
    'max_num_particles': 10,
    'back_relaxation_factor': 1.0,
    'field_output_format': 'paraview',
    'max_frame_index': 100,
    'max_frame_scalar_index': 10000,
    'no_output': 0,
    'no_mesh_output': 0,
    'SA_threshold': 0.0,
    'stat_s_enable': 1,
    'output_factor_x': 1,
    'output_factor_y': 1,
    'output_factor_z': 1,
    'output_x': 1,
    'output_y': 1,
    'output_z': 0,
    'output_velocity_x': 1,
    'output_velocity_y': 1,
    'output_velocity_z': 0,
    'output_pressure': 1,
    'output_density': 0,
    'output_viscosity': 0,
    'output_radiation': 0,
    'output_temperature': 0,
    'output_divergence': 0,
    'output_mass_source_0': 0,
    'output_mass_source_1': 0,
    'output_mass_source_2': 0,
    'output_mass_fraction_0': 0,
    'output_mass_fraction_1': 0,
    'output_mass_fraction_2': 0,
    'output_volume_fraction_0': 1,
    'output_volume_fraction_1': 1,
    'output_volume_fraction_2': 0,
    'output_density_0': 0,
    'output_density_1': 0,
    'output_density_2': 0,
    'output_partial_density_0': 0,
    'output_partial_density_1': 0,
    'output_partial_density_2': 0,
    'output_target_density_0': 0,
    'output_target_density_1': 0,
    'output_target_density_2': 0,
    'output_volume_source': 0,
    'output_curvature': 0,
    'output_excluded': 0,
    'iter_history_enable': 0,
    'iter_history_mesh': 0,
    'iter_history_n': 1,
    'iter_history_sfixed': 0,
}
GENERAL_TYPES = {
    'MODULE': 'string',
    'plt_title': 'string',
    'A': 'vect',
    'B': 'vect',
    'A1': 'vect',
    'B1': 'vect',
    'A2': 'vect',
    'B2': 'vect',
    'box_A': 'vect',
    'box_B': 'vect',
    'IC': 'vect',
    'IR': 'double',
    'IC2': 'vect',
    'IR2': 'double',
    'Nx': 'int',
    'Ny': 'int',
    'Nz': 'int',
    'T': 'double',
    'dt': 'double',
    'dt_auto': 'bool',
    'cfl': 'double',
    'cfl_advection': 'double',
    'num_phases': 'int',
    'gravity': 'vect',
    'force': 'vect',
    'sigma': 'double',
    'density_0': 'double',
    'density_1': 'double',
    'density_2': 'double',
    'viscosity_0': 'double',
    'viscosity_1': 'double',
    'viscosity_2': 'double',
    'molar_0': 'double',
    'molar_1': 'double',
    'molar_2': 'double',
    'deforming_velocity': 'bool',
    'initial_velocity': 'vect',
    'condition_top': 'string',
    'condition_bottom': 'string',
    'condition_left': 'string',
    'condition_right': 'string',
    'condition_close': 'string',
    'condition_far': 'string',
    'chemistry': 'string',
    'chem_intensity': 'double',
    'reaction_zone_lb': 'vect',
    'reaction_zone_rt': 'vect',
    'radiation_enable': 'bool',
    'radiation_intensity': 'double',
    'radiation_direction': 'vect',
    'radiation_box_lb': 'vect',
    'radiation_box_rt': 'vect',
    'absorption_rate_0': 'double',
    'absorption_rate_1': 'double',
    'absorption_rate_2': 'double',
    'heat_enable': 'bool',
    'linear_solver_heat': 'string',
    'heat_box_lb': 'vect',
    'heat_box_rt': 'vect',
    'conductivity_0': 'double',
    'conductivity_1': 'double',
    'conductivity_2': 'double',
    'temperature_initial': 'double',
    'heat_box_temperature': 'double',
    'heat_relaxation_factor': 'double',
    'time_second_order_heat': 'bool',
    'incompressible_relaxation': 'double',
    'temperature_expansion_rate_0': 'double',
    'temperature_expansion_rate_1': 'double',
    'temperature_expansion_rate_2': 'double',
    'temperature_expansion_base_0': 'double',
    'temperature_expansion_base_1': 'double',
    'temperature_expansion_base_2': 'double',
    'fluid_enable': 'bool',
    'advection_enable': 'bool',
    'advection_solver': 'string',
    'advection_dt_factor': 'double',
    'tvd_split': 'bool',
    'convergence_tolerance': 'double',
    'num_iterations_limit': 'int',
    'velocity_relaxation_factor': 'double',
    'pressure_relaxation_factor': 'double',
    'linear_solver_velocity': 'string',
    'linear_solver_pressure': 'string',
    'lu_relaxed_relaxation_factor': 'double',
    'lu_relaxed_num_iters_limit': 'int',
    'lu_relaxed_tolerance': 'double',
    'time_second_order': 'bool',
    'rhie_chow_factor': 'double',
    'simpler': 'bool',
    'initial_volume_fraction_smooth_times': 'int',
    'density_smooth_times': 'int',
    'viscosity_smooth_times': 'int',
    'force_smooth_times': 'int',
    'force_geometric_average': 'bool',
    'guess_extrapolation': 'double',
    'compressible_enable': 'bool',
    'meshvel': 'vect',
    'meshvel_output': 'bool',
    'meshvel_weight': 'double',
    'sharp': 'double',
    'spawning_gap': 'double',
    'particle_radius': 'double',
    'min_num_particles': 'int',
    'max_num_particles': 'int',
    'back_relaxation_factor': 'double',
    'field_output_format': 'string',
    'max_frame_index': 'int',
    'max_frame_scalar_index': 'int',
    'no_output': 'bool',
    'no_mesh_output': 'bool',
    'SA_threshold': 'double',
    'stat_s_enable': 'bool',
    'output_factor_x': 'int',
    'output_factor_y': 'int',
    'output_factor_z': 'int',
    'output_x': 'bool',
    'output_y': 'bool',
    'output_z': 'bool',
    'output_velocity_x': 'bool',
    'output_velocity_y': 'bool',
    'output_velocity_z': 'bool',
    'output_pressure': 'bool',
    'output_density': 'bool',
    'output_viscosity': 'bool',
    'output_radiation': 'bool',
    'output_temperature': 'bool',
    'output_divergence': 'bool',
    'output_mass_source_0': 'bool',
    'output_mass_source_1': 'bool',
    'output_mass_source_2': 'bool',
    'output_mass_fraction_0': 'bool',
    'output_mass_fraction_1': 'bool',
    'output_mass_fraction_2': 'bool',
    'output_volume_fraction_0': 'bool',
    'output_volume_fraction_1': 'bool',
    'output_volume_fraction_2': 'bool',
    'output_density_0': 'bool',
    'output_density_1': 'bool',
    'output_density_2': 'bool',
    'output_partial_density_0': 'bool',
    'output_partial_density_1': 'bool',
    'output_partial_density_2': 'bool',
    'output_target_density_0': 'bool',
    'output_target_density_1': 'bool',
    'output_target_density_2': 'bool',
    'output_volume_source': 'bool',
    'output_curvature': 'bool',
    'output_excluded': 'bool',
    'iter_history_enable': 'bool',
    'iter_history_mesh': 'bool',
    'iter_history_n': 'int',
    'iter_history_sfixed': 'int',
}

MODULE_DIM = {"hydro2D_uniform_MPI": 2, "hydro2d": 2, "hydro3D_uniform_MPI": 3, "hydro3d": 3,
              # aliases registered by the GPU module
              "hydro2d_gpu": 2, "hydro3d_gpu": 3}


def _vec3(v):
    v = list(v) + [0.0] * 3
    return [float(x) for x in v[:3]]


class Params(dict):
    """Key-value parameters with the reference's names; missing keys raise like P_x["k"]
    (common/data_structures.hpp:238-242)."""

    def __init__(self, *overrides, **kw):
        super().__init__(GENERAL_DEFAULTS)
        for o in overrides:
            self.update(o)
        self.update(kw)

    # -- .hydroconf text -----------------------------------------------------
    _set_re = re.compile(r"^\s*set\s+(\w+)\s+(\w+)\s+(.*?)\s*$")

    @staticmethod
    def _as_console_text(v):
        """A parameter value as `$(name)` substitutes it: written with the stream's default formatting (6 significant digits,
        `2.0` -> `2`) and read back as one token (console.cpp:464-509), so `$(T)0` with T = 2 is `20`."""
        if isinstance(v, bool):
            return "1" if v else "0"
        if isinstance(v, float):
            return "%g" % v
        if isinstance(v, (tuple, list)):
            return "(" + ",".join("%g" % x for x in v) + ")"
        return str(v).split()[0] if str(v).split() else ""

    def read_hydroconf(self, text):
        """Apply `set <type> <name> <value>` / `del <name>` lines (console.cpp:276-314)."""
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if line.startswith("del "):
                self.pop(line.split()[1], None)
                continue
            m = self._set_re.match(line)
            if not m:
                continue  # console commands (ae, run, init, start ...) are the caller's business
            typ, name, val = m.groups()
            val = re.sub(r"\$\((\w+)\)|\$(\w+)", lambda g: self._as_console_text(self[g.group(1) or g.group(2)]), val)
            if typ in ("int", "c_int"):
                self[name] = int(val)
            elif typ in ("double", "c_double"):
                self[name] = float(val)
            elif typ in ("bool", "c_bool"):
                self[name] = int(val)
            elif typ == "vect":
                self[name] = tuple(float(x) for x in val.strip("() ").replace(",", " ").split())
            else:
                self[name] = val.strip('"')
        return self

    # -- struct -----------------------------------------------------------------
    def to_struct(self, world_size=1, rank=0, device=0, solver_ctas=0):
        p = self
        reject_unsupported(p)

        def need(k):
            if k not in p:
                raise KeyError("'%s' undefined" % k)
            return p[k]

        c = HgConfig()
        module = need("MODULE")
        if module not in MODULE_DIM:
            raise ValueError("Unknown module '%s'" % module)
        c.dim = MODULE_DIM[module]
        c.Nx, c.Ny, c.Nz = int(need("Nx")), int(need("Ny")), int(need("Nz"))  # Nz ignored for dim == 2
        for name in ("A", "B", "box_A", "box_B", "A1", "B1", "A2", "B2", "IC", "IC2", "gravity", "force",
                     "meshvel", "heat_box_lb", "heat_box_rt"):
            setattr(c, name, d3(*_vec3(need(name))))
        c.IR, c.IR2 = float(need("IR")), float(need("IR2"))
        for i, side in enumerate(SIDES):
            words = str(need("condition_" + side)).split()
            if words[0] not in BC_KINDS:
                raise ValueError("Parse: Unknown boundary condition type")
            c.condition_kind[i] = BC_KINDS[words[0]]
            vel = [float(w) for w in words[1:1 + c.dim]]
            c.condition_velocity[i] = d3(*_vec3(vel))
        if "pressure_fixed_point" in p:
            c.pressure_fixed_enable = 1
            c.pressure_fixed_point = d3(*_vec3(p["pressure_fixed_point"]))
            c.pressure_fixed_value = float(p.get("pressure_fixed_value", 0.0))
        c.initial_velocity = d3(*_vec3(p.get("initial_velocity", (0, 0, 0))))
        c.initial_pois = int(bool(p.get("initial_pois", 0)))
        if "initial_sin_n" in p:
            c.initial_sin_enable = 1
            c.initial_sin_n = d3(*_vec3(p["initial_sin_n"]))
            c.initial_sin_lambda = float(need("initial_sin_lambda"))
            c.initial_sin_phase = float(need("initial_sin_phase"))
        c.num_phases = int(need("num_phases"))
        for i in range(HG_MAX_PHASES):
            if i < c.num_phases:
                need("density_%d" % i), need("viscosity_%d" % i), need("conductivity_%d" % i)
            c.density[i] = float(p.get("density_%d" % i, 1.0))
            c.viscosity[i] = float(p.get("viscosity_%d" % i, 1.0))
            c.conductivity[i] = float(p.get("conductivity_%d" % i, 1.0))
            c.initial_volume_fraction[i] = float(p.get("initial_volume_fraction_%d" % i, 0.0))
            c.enable_settling[i] = int(bool(int(p.get("enable_settling_%d" % i, 0))))
            c.bubble_radius[i] = float(p.get("bubble_radius_%d" % i, 0.0)) if c.enable_settling[i] else 0.0
        for name in ("initial_volume_fraction_smooth_times", "dt_auto", "fluid_enable", "advection_enable",
                     "num_iterations_limit", "time_second_order", "simpler", "force_geometric_average",
                     "lu_relaxed_num_iters_limit", "density_smooth_times", "viscosity_smooth_times",
                     "force_smooth_times", "tvd_split", "heat_enable", "time_second_order_heat", "meshvel_output"):
            setattr(c, name, int(need(name)))
        for name in ("dt", "cfl", "cfl_advection", "sigma", "convergence_tolerance", "velocity_relaxation_factor",
                     "pressure_relaxation_factor", "rhie_chow_factor", "guess_extrapolation",
                     "lu_relaxed_tolerance", "lu_relaxed_relaxation_factor", "advection_dt_factor", "sharp",
                     "temperature_initial", "heat_box_temperature", "heat_relaxation_factor"):
            setattr(c, name, float(need(name)))
        for name in ("linear_solver_velocity", "linear_solver_pressure", "linear_solver_heat"):
            v = need(name)
            if v not in LINEAR_SOLVERS:
                raise ValueError("Unknown linear solver '%s'" % v)
            setattr(c, name, LINEAR_SOLVERS[v])
        # automatic mesh velocity (hydro2d.hpp:1510-1524): present = on; the reference asserts on any other value
        if "meshvel_auto" in p:
            if p["meshvel_auto"] not in MESHVEL_AUTO:
                raise ValueError("Unknown meshvel_auto=%s" % p["meshvel_auto"])
            c.meshvel_auto = MESHVEL_AUTO[p["meshvel_auto"]]
            c.meshvel_weight = float(need("meshvel_weight"))
        else:
            c.meshvel_auto, c.meshvel_weight = 0, float(p.get("meshvel_weight", 0.5))
        if need("advection_solver") != "tvd":
            raise ValueError("only advection_solver tvd is on the GPU path")
        c.world_size, c.rank, c.device = world_size, rank, device
        c.pressure_sweeps_per_check = int(p.get("pressure_sweeps_per_check", 0))
        c.solver_ctas = int(solver_ctas)
        return c

    def hydroconf_lines(self):
        """The same parameters as `set` lines for the reference binary."""
        out = []
        for k, v in self.items():
            if k == "pressure_sweeps_per_check":
                continue
            if isinstance(v, str):
                out.append('set string %s "%s"' % (k, v) if " " in v else "set string %s %s" % (k, v))
            elif isinstance(v, (tuple, list)):
                out.append("set vect %s (%s)" % (k, ", ".join(repr(float(x)) for x in v)))
            elif isinstance(v, float) or k in _DOUBLE_KEYS or GENERAL_TYPES.get(k) == "double":
                out.append("set double %s %r" % (k, float(v)))
            else:
                typ = "bool" if (k in _BOOL_KEYS or GENERAL_TYPES.get(k) == "bool") else "int"
                out.append("set %s %s %d" % (typ, k, int(v)))
        return out


def reject_unsupported(p):
    """Options that change the reference's results but are not on the GPU path fail loudly instead of being
    dropped (hydro2d.hpp:326-368 initial images / deforming velocity, 1030-1122 phase slip, 1129/1387 compressibility,
    1163-1185 chemistry, 1294 radiation)."""
    def truthy(k):
        v = p.get(k, 0)
        return bool(int(v)) if not isinstance(v, str) else v not in ("", "0")
    for k in ("compressible_enable", "deforming_velocity", "radiation_enable"):
        if truthy(k):
            raise ValueError("%s 1 is not on the GPU path" % k)
    if str(p.get("chemistry", "steady")) != "steady" or float(p.get("chem_intensity", 0.0)) != 0.0:
        raise ValueError("chemistry other than 'steady' with chem_intensity 0 is not on the GPU path")
    for k in ("imgu_init", "imgv_init", "img_init"):
        if k in p:
            raise ValueError("%s is not on the GPU path" % k)
    if truthy("velocity_is_carrier"):
        raise ValueError("velocity_is_carrier 1 (mixture volume source) is not on the GPU path")
    if float(p.get("antidiffusion_factor", 0.0)) != 0.0:
        raise ValueError("antidiffusion_factor != 0 is not on the GPU path")


_DOUBLE_KEYS = {"bubble_radius_0", "bubble_radius_1", "bubble_radius_2", "antidiffusion_factor", "initial_sin_lambda", "initial_sin_phase", "pressure_fixed_value", "T", "dt",
                "initial_volume_fraction_0", "initial_volume_fraction_1", "initial_volume_fraction_2"}
_BOOL_KEYS = {"enable_settling_0", "enable_settling_1", "enable_settling_2", "velocity_is_carrier", "dt_auto", "deforming_velocity", "radiation_enable", "heat_enable", "time_second_order_heat",
              "fluid_enable", "advection_enable", "tvd_split", "time_second_order", "simpler",
              "force_geometric_average", "compressible_enable", "initial_pois", "no_output", "no_mesh_output"}
