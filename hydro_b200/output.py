"""Mesh output in the reference's ParaView format.

`write_vts` reproduces `output::SessionParaviewStructured::CreateDataFile`
(source/hydro2dmpi/output_paraview.hpp:34-173): ASCII StructuredGrid, arrays labelled Float32, values printed
with the stream's default 6 significant digits (`out << value << " "`), PointData x/y/z, CellData in the order
of the content pool (hydro2d.hpp:711-818) filtered by the `output_<name>` switches, then the node coordinates.
Fields are fetched from the device only here (hg_get_field), i.e. at frame times, not every step.
`write_pvd` writes the `.pvd` collection (output_paraview.hpp:133-150).  output_factor_* = 1 only.
"""
import numpy as np


def _fmt(a):
    return " ".join("%g" % v for v in a) + " "


def _nodes(params, dim):
    n = [int(params["Nx"]), int(params["Ny"]), int(params["Nz"]) if dim == 3 else 0]
    A = list(params["A"]) + [0.0] * 3
    B = list(params["B"]) + [0.0] * 3
    ax = []
    for d in range(3):
        if d < dim:
            # InitUniformMesh (mesh.hpp:722-736): lb + (midx / mesh_size) * (rt - lb)
            ax.append(np.array([A[d] + (i / n[d]) * (B[d] - A[d]) for i in range(n[d] + 1)]))
        else:
            ax.append(np.array([0.0]))
    return n, ax


def cell_content(hydro, params):
    """(name, values) pairs of the selected cell entries, in content-pool order."""
    dim = hydro.dim
    p = params
    out = []

    def on(name):
        return bool(p.get("output_" + name, 0))

    if on("velocity_x"):
        out.append(("velocity_x", hydro.get("VELOCITY_X")))
    if on("velocity_y"):
        out.append(("velocity_y", hydro.get("VELOCITY_Y")))
    if on("velocity_z"):
        out.append(("velocity_z", hydro.get("VELOCITY_Z") if dim == 3 else np.zeros(hydro.nc)))
    if on("pressure"):
        out.append(("pressure", hydro.get("PRESSURE")))
    if on("viscosity"):
        out.append(("viscosity", hydro.get("VISCOSITY")))
    if on("temperature"):
        out.append(("temperature", hydro.get("TEMPERATURE")))
    if on("excluded"):
        out.append(("excluded", hydro.get("EXCLUDED")))
    np_ = int(p["num_phases"])
    for i in range(np_):
        if on("partial_density_%d" % i):
            out.append(("partial_density_%d" % i, hydro.get("PARTIAL_DENSITY_%d" % i)))
    for i in range(np_):
        if on("volume_fraction_%d" % i):
            out.append(("volume_fraction_%d" % i, hydro.get("VOLUME_FRACTION_%d" % i)))
    return out


def vts_text(hydro, params):
    dim = hydro.dim
    n, ax = _nodes(params, dim)
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    order = lambda a: a.transpose(2, 1, 0).reshape(-1)   # x fastest (mesh.hpp:552-561)
    xs, ys, zs = order(X), order(Y), order(Z)
    ext = "0 %d 0 %d 0 %d" % (n[0], n[1], n[2] if dim == 3 else 0)
    L = ['<?xml version="1.0"?>',
         '<VTKFile type="StructuredGrid" version="0.1" byte_order="LittleEndian">',
         '  <StructuredGrid WholeExtent="%s">' % ext, '    <Piece Extent="%s">' % ext, "      <PointData>"]

    def arr(name, ncomp, vals):
        L.append('        <DataArray Name="%s" NumberOfComponents="%d" type="Float32" format="ascii">' % (name, ncomp))
        L.append(_fmt(vals))
        L.append("        </DataArray>")

    for name, vals in (("x", xs), ("y", ys), ("z", zs if dim == 3 else np.zeros_like(xs))):
        if params.get("output_" + name, 0):
            arr(name, 1, vals)
    L.append("      </PointData>")
    L.append("      <CellData>")
    for name, vals in cell_content(hydro, params):
        arr(name, 1, vals)
    L.append("      </CellData>")
    L.append("      <Points>")
    L.append('        <DataArray Name="mesh" NumberOfComponents="3" type="Float32" format="ascii">')
    pts = np.stack([xs, ys, zs if dim == 3 else np.zeros_like(xs)], axis=1).reshape(-1)
    L.append(_fmt(pts) + "        </DataArray>")   # the reference writes no newline after the node list
    L.append("      </Points>")
    L += ["    </Piece>", "  </StructuredGrid>", "</VTKFile>"]
    return "\n".join(L) + "\n"


def write_vts(hydro, params, path):
    with open(path, "w") as f:
        f.write(vts_text(hydro, params))


def write_pvd(entries, path):
    """entries: list of (time, datafile name)."""
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian">\n  <Collection>\n')
        for t, name in entries:
            f.write('    <DataSet timestep="%g" group="" part="0" file="%s"/>\n' % (t, name))
        f.write("  </Collection>\n</VTKFile>\n")
