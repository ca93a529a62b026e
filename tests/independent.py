"""From-scratch plane-wave solutions of the layered elastic half-space (NumPy only), used as
INDEPENDENT checks of the oracle: nothing here follows the reference's formulation (Haskell / Dunkin
propagators); every layer simply carries four potentials and all boundary conditions are assembled
into one linear system."""
import numpy as np


def layer_system(w, p, thk, vp, vs, rho, incident="P"):
    """Boundary-condition matrix M, right-hand side for a unit up-going P (or SV) wave in the half-space, and
    the four solution columns (ux, uz, tzz, txz) of the top layer at z = 0.
    Plane waves exp(i w (p x - t)), z down.  Unknowns: (P down, P up, S down, S up) per finite layer,
    (P down, S down) in the half-space.  Rows: tzz = txz = 0 at the surface, continuity of
    (ux, uz, tzz, txz) at every interface.  thk[-1] is ignored (half-space)."""
    n = len(thk)
    k = w * p

    def cols(a, b, r, z):
        nua = np.sqrt(complex((w / a)**2 - k * k))
        nub = np.sqrt(complex((w / b)**2 - k * k))
        if nua.imag < 0:
            nua = -nua
        if nub.imag < 0:
            nub = -nub
        mu = r * b * b
        lam = r * a * a - 2 * mu
        out = []
        for s, nu, kind in ((+1, nua, 'p'), (-1, nua, 'p'), (+1, nub, 's'), (-1, nub, 's')):
            e = np.exp(1j * s * nu * z)
            kz = s * nu
            if kind == 'p':   # u = grad(phi), phi = exp(i (k x + kz z))
                ux, uz = 1j * k, 1j * kz
            else:             # u = curl(psi y): ux = -dpsi/dz, uz = dpsi/dx
                ux, uz = -1j * kz, 1j * k
            dux_dz, duz_dz, dux_dx, duz_dx = 1j * kz * ux, 1j * kz * uz, 1j * k * ux, 1j * k * uz
            tzz = lam * (dux_dx + duz_dz) + 2 * mu * duz_dz
            txz = mu * (dux_dz + duz_dx)
            out.append(np.array([ux, uz, tzz, txz]) * e)
        return out
    N = 4 * (n - 1) + 2
    M = np.zeros((N, N), dtype=complex)
    rhs = np.zeros(N, dtype=complex)
    c0 = cols(vp[0], vs[0], rho[0], 0.0)
    for j in range(4):
        M[0, j] = c0[j][2]
        M[1, j] = c0[j][3]
    row = 2
    for m in range(n - 1):
        cb = cols(vp[m], vs[m], rho[m], thk[m])
        for j in range(4):
            M[row:row + 4, 4 * m + j] = cb[j]
        if m + 1 < n - 1:
            ct = cols(vp[m + 1], vs[m + 1], rho[m + 1], 0.0)
            for j in range(4):
                M[row:row + 4, 4 * (m + 1) + j] = -ct[j]
        else:
            ch = cols(vp[n - 1], vs[n - 1], rho[n - 1], 0.0)
            M[row:row + 4, 4 * (n - 1) + 0] = -ch[0]   # P leaving downwards
            M[row:row + 4, 4 * (n - 1) + 1] = -ch[2]   # S leaving downwards
            rhs[row:row + 4] = ch[1] if incident == "P" else ch[3]   # the incident wave, amplitude 1
        row += 4
    return M, rhs, c0


def surface_response(w, p, thk, vp, vs, rho, incident="P"):
    """(ux, uz) at the free surface for a unit P (or SV) wave incident from the half-space."""
    M, rhs, c0 = layer_system(w, p, thk, vp, vs, rho, incident)
    sc = np.max(np.abs(M), axis=0)
    sc[sc == 0] = 1
    x = np.linalg.solve(M / sc, rhs) / sc
    u = sum(x[j] * c0[j][:2] for j in range(4))
    return u[0], u[1]


def rayleigh_secular(c, T, thk, vp, vs, rho):
    """Determinant of the homogeneous system (no incident wave) at phase velocity c and period T,
    columns scaled to unit max-norm (positive factors: the phase of the determinant is untouched).
    Free Rayleigh modes are its zeros for c below the half-space S velocity."""
    M, _, _ = layer_system(2 * np.pi / T, 1.0 / c, thk, vp, vs, rho)
    sc = np.max(np.abs(M), axis=0)
    sc[sc == 0] = 1
    return np.linalg.det(M / sc)


def love_secular(c, T, thk, vs, rho):
    """SH analogue of rayleigh_secular: two plane waves per finite layer (down, up), one decaying wave
    in the half-space; tau_yz = 0 at the surface, (u_y, tau_yz) continuous at the interfaces."""
    n = len(thk)
    w = 2 * np.pi / T
    k = w / c

    def cols(b, r, z):
        nu = np.sqrt(complex((w / b)**2 - k * k))
        if nu.imag < 0:
            nu = -nu
        mu = r * b * b
        return [np.array([1.0, mu * 1j * s * nu]) * np.exp(1j * s * nu * z) for s in (+1, -1)]
    N = 2 * (n - 1) + 1
    M = np.zeros((N, N), dtype=complex)
    c0 = cols(vs[0], rho[0], 0.0)
    M[0, 0], M[0, 1] = c0[0][1], c0[1][1]
    row = 1
    for m in range(n - 1):
        cb = cols(vs[m], rho[m], thk[m])
        M[row:row + 2, 2 * m] = cb[0]
        M[row:row + 2, 2 * m + 1] = cb[1]
        if m + 1 < n - 1:
            ct = cols(vs[m + 1], rho[m + 1], 0.0)
            M[row:row + 2, 2 * (m + 1)] = -ct[0]
            M[row:row + 2, 2 * (m + 1) + 1] = -ct[1]
        else:
            M[row:row + 2, 2 * (n - 1)] = -cols(vs[n - 1], rho[n - 1], 0.0)[0]
        row += 2
    sc = np.max(np.abs(M), axis=0)
    sc[sc == 0] = 1
    return np.linalg.det(M / sc)


def rayleigh_secular_ocean(c, T, thk, vp, vs, rho):
    """rayleigh_secular for a model whose TOP layer is a fluid (vs[0] == 0): the water carries two
    compressional waves; pressure-free surface; at the sea floor u_z and tau_zz are continuous and the
    shear traction of the solid vanishes."""
    n = len(thk)
    w = 2 * np.pi / T
    k = w / c

    def nu_of(v):
        nu = np.sqrt(complex((w / v)**2 - k * k))
        return -nu if nu.imag < 0 else nu

    def fluid_cols(a, r, z):
        nua = nu_of(a)
        lam = r * a * a
        out = []
        for s in (+1, -1):
            kz = s * nua
            ux, uz = 1j * k, 1j * kz
            tzz = lam * (1j * k * ux + 1j * kz * uz)
            out.append(np.array([ux, uz, tzz, 0.0]) * np.exp(1j * kz * z))
        return out

    def solid_cols(a, b, r, z):
        nua, nub = nu_of(a), nu_of(b)
        mu = r * b * b
        lam = r * a * a - 2 * mu
        out = []
        for s, nu, kind in ((+1, nua, 'p'), (-1, nua, 'p'), (+1, nub, 's'), (-1, nub, 's')):
            kz = s * nu
            ux, uz = (1j * k, 1j * kz) if kind == 'p' else (-1j * kz, 1j * k)
            dux_dz, duz_dz, dux_dx, duz_dx = 1j * kz * ux, 1j * kz * uz, 1j * k * ux, 1j * k * uz
            out.append(np.array([ux, uz, lam * (dux_dx + duz_dz) + 2 * mu * duz_dz, mu * (dux_dz + duz_dx)])
                       * np.exp(1j * kz * z))
        return out
    N = 2 + 4 * (n - 2) + 2
    M = np.zeros((N, N), dtype=complex)
    f0, fb = fluid_cols(vp[0], rho[0], 0.0), fluid_cols(vp[0], rho[0], thk[0])
    M[0, 0], M[0, 1] = f0[0][2], f0[1][2]                      # pressure-free sea surface

    def place(row, col0, cs, sign, rows):
        for j, cvec in enumerate(cs):
            M[row:row + len(rows), col0 + j] = sign * cvec[rows]
    row = 1
    col = 2
    # sea floor: uz, tzz continuous; txz of the solid = 0
    top = solid_cols(vp[1], vs[1], rho[1], 0.0) if n > 2 else None
    if n > 2:
        place(row, 0, fb, +1, [1, 2])
        place(row, col, top, -1, [1, 2])
        for j in range(4):
            M[row + 2, col + j] = top[j][3]
        row += 3
        for m in range(1, n - 1):
            cb = solid_cols(vp[m], vs[m], rho[m], thk[m])
            place(row, col, cb, +1, [0, 1, 2, 3])
            if m + 1 < n - 1:
                place(row, col + 4, solid_cols(vp[m + 1], vs[m + 1], rho[m + 1], 0.0), -1, [0, 1, 2, 3])
            else:
                ch = solid_cols(vp[n - 1], vs[n - 1], rho[n - 1], 0.0)
                M[row:row + 4, col + 4] = -ch[0]
                M[row:row + 4, col + 5] = -ch[2]
            row += 4
            col += 4
    sc = np.max(np.abs(M), axis=0)
    sc[sc == 0] = 1
    return np.linalg.det(M / sc)
