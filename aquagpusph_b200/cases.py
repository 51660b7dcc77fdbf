"""Synthetic particle sets shared by the CPU and GPU tests and by bench.py.

`dam_break(dims, n_side, ...)` builds a small version of the SPHERIC dam-break
geometry of the reference examples (examples/{2D,3D}/spheric_testcase*_dambreak):
a jittered fluid lattice (imove = 1) resting on boundary-integral elements
(imove = -3, normals pointing out of the fluid, m = element area), a couple of
sensors (imove = 0) and a few buffer particles (imove = -255) parked at
domain_max, with a hydrostatic-ish density field and random velocities.
"""
import math

import numpy as np


def vs(dims):
    return 4 if dims == 3 else 2


def dam_break(dims=3, n_side=12, hfac=2.0, seed=1234, jitter=0.2, n_sensors=3, n_buffer=5,
              shuffle=True):
    rng = np.random.default_rng(seed)
    dr = np.float32(0.01)
    h = np.float32(hfac * dr)
    cs, refd = np.float32(40.0), np.float32(998.0)
    V = vs(dims)
    # fluid block
    ax = [np.arange(n_side) for _ in range(dims)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, dims).astype(np.float64)
    rf = (g + 0.5) * dr + jitter * dr * rng.uniform(-1, 1, g.shape)
    nf = rf.shape[0]
    L = n_side * dr
    # boundary elements: floor (last axis = 0 plane) and one wall (axis 0 = 0 plane)
    bax = [np.arange(-2, n_side + 2) for _ in range(dims - 1)]
    bg = np.stack(np.meshgrid(*bax, indexing="ij"), -1).reshape(-1, dims - 1).astype(np.float64)
    nb1 = bg.shape[0]
    floor = np.zeros((nb1, dims))
    floor[:, :dims - 1] = (bg + 0.5) * dr
    nfloor = np.zeros((nb1, dims)); nfloor[:, dims - 1] = -1.0
    wall = np.zeros((nb1, dims))
    wall[:, 1:] = (bg + 0.5) * dr
    nwall = np.zeros((nb1, dims)); nwall[:, 0] = -1.0
    rb = np.concatenate([floor, wall]); nrm = np.concatenate([nfloor, nwall])
    nb = rb.shape[0]
    # sensors inside the fluid, buffer particles far away
    rs = rng.uniform(0.2 * L, 0.8 * L, (n_sensors, dims))
    domain_min = np.full(dims, -0.5 * L - 0.1)
    domain_max = np.full(dims, 2.0 * L + 0.1)
    rbuf = np.tile(domain_max, (n_buffer, 1))

    N = nf + nb + n_sensors + n_buffer
    r = np.zeros((N, V), np.float32)
    r[:, :dims] = np.concatenate([rf, rb, rs, rbuf]).astype(np.float32)
    imove = np.concatenate([np.ones(nf), -3 * np.ones(nb), np.zeros(n_sensors),
                            -255 * np.ones(n_buffer)]).astype(np.int32)
    iset = np.concatenate([np.zeros(nf), np.ones(nb), 2 * np.ones(n_sensors),
                           np.zeros(n_buffer)]).astype(np.uint32)
    normal = np.zeros((N, V), np.float32)
    normal[nf:nf + nb, :dims] = nrm
    tangent = np.zeros((N, V), np.float32)
    tangent[nf:nf + nb, (1 if dims == 3 else 0)] = 1.0
    depth = np.clip(L - r[:, dims - 1], 0, None)
    rho = (refd * (1.0 + 9.81 * depth / cs ** 2) *
           (1.0 + 1e-3 * rng.uniform(-1, 1, N))).astype(np.float32)
    rho[imove == -255] = refd
    m = np.zeros(N, np.float32)
    m[:nf] = refd * dr ** dims
    m[nf:nf + nb] = dr ** (dims - 1)            # element area (length in 2-D)
    m[nf + nb:nf + nb + n_sensors] = refd * dr ** dims
    u = np.zeros((N, V), np.float32)
    u[:nf, :dims] = 0.5 * rng.uniform(-1, 1, (nf, dims))
    dudt = np.zeros((N, V), np.float32)
    dudt[:nf, :dims] = 5.0 * rng.uniform(-1, 1, (nf, dims))
    drhodt = np.zeros(N, np.float32)
    drhodt[:nf] = 10.0 * rng.uniform(-1, 1, nf)

    if shuffle:
        perm = rng.permutation(N)
        r, imove, iset, normal, tangent, rho, m, u, dudt, drhodt = (
            a[perm] for a in (r, imove, iset, normal, tangent, rho, m, u, dudt, drhodt))

    g = np.zeros(V, np.float32); g[dims - 1] = -9.81
    dmin = np.zeros(V, np.float32); dmin[:dims] = domain_min
    dmax = np.zeros(V, np.float32); dmax[:dims] = domain_max
    return dict(
        dims=dims, N=N, h=float(h), dr=float(dr), cs=float(cs), p0=0.0, support=2.0,
        refd=np.array([refd, refd, refd], np.float32),
        visc_dyn=np.array([1e-3, 0.0, 0.0], np.float32),
        delta=np.array([0.1, 0.0, 0.0], np.float32),
        g=g, domain_min=dmin, domain_max=dmax, courant=0.25, dt_Ma=0.1, dt_min=1e-7,
        id=np.arange(N, dtype=np.uint32), r=r, imove=imove, iset=iset, normal=normal,
        tangent=tangent, rho=rho, m=m, u=u, dudt=dudt, drhodt=drhodt,
    )


def lattice(n_side, hfac=2.0, dims=3, seed=1234, cs=40.0, uscale=0.01, jitter=0.0):
    """BASELINE config 5 shape: uniform lattice dr = 1, all fluid, rho = refd,
    u = 0.01*cs*U(-1,1); no boundary (the ref has no periodic BC).  `uscale` / `jitter` (tests):
    faster particles on off-lattice positions, so that they cross the cuts of a slab run."""
    rng = np.random.default_rng(seed)
    V = vs(dims)
    ax = [np.arange(n_side, dtype=np.float32) for _ in range(dims)]
    gpos = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, dims)
    N = gpos.shape[0]
    r = np.zeros((N, V), np.float32)
    r[:, :dims] = gpos + 0.5
    u = np.zeros((N, V), np.float32)
    u[:, :dims] = (uscale * cs * rng.uniform(-1, 1, (N, dims))).astype(np.float32)
    if jitter:
        r[:, :dims] += (jitter * np.random.default_rng(seed + 1).uniform(-1, 1, (N, dims))).astype(np.float32)
    refd = np.float32(1000.0)
    z = np.zeros(V, np.float32)
    return dict(
        dims=dims, N=N, h=float(hfac), dr=1.0, cs=float(cs), p0=0.0, support=2.0,
        refd=np.array([refd], np.float32), visc_dyn=np.array([1e-3], np.float32),
        delta=np.array([0.1], np.float32), g=z.copy(), domain_min=z.copy() - 10.0,
        domain_max=z.copy() + n_side + 10.0, courant=0.25, dt_Ma=0.1, dt_min=1e-7,
        id=np.arange(N, dtype=np.uint32), r=r, imove=np.ones(N, np.int32),
        iset=np.zeros(N, np.uint32), normal=np.zeros((N, V), np.float32),
        tangent=np.zeros((N, V), np.float32),
        rho=(refd * (1.0 + 1e-3 * rng.uniform(-1, 1, N))).astype(np.float32),
        m=np.full(N, refd, np.float32), u=u, dudt=np.zeros((N, V), np.float32),
        drhodt=np.zeros(N, np.float32),
    )


def spheric2_dam_break(n=1000000, hfac=3.0, seed=None, jitter=0.0, uscale=0.1):
    """BASELINE config 2: the 3-D SPHERIC test 2 dam break with obstacle, same
    geometry and field initialisation as the reference's case generator
    (examples/3D/spheric_testcase2_dambreak/src/Create.py:39-462): fluid block
    l x d x h = 1.228 x 1.0 x 0.55 m, box obstacle, closed tank, boundary-integral
    elements (imove = -3, m = dr^2) with outward normals, 8 pressure sensors.
    `n` is the requested number of FLUID particles (Create.py:62)."""
    g, cs, courant, refd = 9.81, 40.0, 0.25, 998.0
    delta, visc_dyn = 1.0, 0.000894
    h, l, d = 0.55, 1.228, 1.0
    H_box, L_box, D_box, x_box = 0.161, 0.161, 0.403, -1.248
    H, L, D = 1.0, l + abs(x_box) + 0.744, d
    dr = (l * d * h / n) ** (1.0 / 3.0)
    nx, ny, nz = int(round(l / dr)), int(round(d / dr)), int(round(h / dr))
    hFluid = nz * dr
    Nx_box, Ny_box, Nz_box = (int(round(v / dr)) for v in (L_box, D_box, H_box))
    x_box = (int(round(x_box / dr - 0.5 * Nx_box)) + 0.5 * Nx_box) * dr
    Nx, Ny, Nz = int(round(L / dr)), int(round(D / dr)), int(round(H / dr))

    def grid(*axes):
        return np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, len(axes))

    parts = []  # (pos[n,3], normal, tangent, imove)

    def face(pos, normal, tangent):
        parts.append((pos, np.tile(normal, (len(pos), 1)), np.tile(tangent, (len(pos), 1)),
                      np.full(len(pos), -3, np.int32)))

    # fluid (x outer, y, z inner like Create.py:107-134)
    gi = grid(np.arange(nx), np.arange(ny), np.arange(nz)).astype(np.float64)
    pf = np.stack([(gi[:, 0] + 0.5) * dr, -0.5 * Ny * dr + (gi[:, 1] + 0.5) * dr,
                   (gi[:, 2] + 0.5) * dr], 1)
    nfluid = len(pf)
    press = refd * g * (hFluid - pf[:, 2]) * np.cos(0.5 * np.pi * (l - pf[:, 0]) / l)
    dens_f = refd + press / cs ** 2
    parts.append((pf, np.zeros((nfluid, 3)), np.zeros((nfluid, 3)), np.ones(nfluid, np.int32)))
    # box obstacle (Create.py:136-252)
    x0, y0, z0 = x_box - 0.5 * Nx_box * dr, -0.5 * Ny_box * dr, 0.0
    ix, iy, iz = np.arange(Nx_box), np.arange(Ny_box), np.arange(Nz_box)
    q = grid(ix, iy).astype(np.float64)
    face(np.stack([x0 + (q[:, 0] + .5) * dr, y0 + (q[:, 1] + .5) * dr,
                   np.full(len(q), z0 + Nz_box * dr)], 1), (0, 0, 1), (-1, 0, 0))
    q = grid(iy, iz).astype(np.float64)
    face(np.stack([np.full(len(q), x0), y0 + (q[:, 0] + .5) * dr, z0 + (q[:, 1] + .5) * dr], 1),
         (1, 0, 0), (0, -1, 0))
    face(np.stack([np.full(len(q), x0 + Nx_box * dr), y0 + (q[:, 0] + .5) * dr,
                   z0 + (q[:, 1] + .5) * dr], 1), (-1, 0, 0), (0, -1, 0))
    q = grid(ix, iz).astype(np.float64)
    face(np.stack([x0 + (q[:, 0] + .5) * dr, np.full(len(q), y0), z0 + (q[:, 1] + .5) * dr], 1),
         (0, 1, 0), (1, 0, 0))
    face(np.stack([x0 + (q[:, 0] + .5) * dr, np.full(len(q), y0 + Ny_box * dr),
                   z0 + (q[:, 1] + .5) * dr], 1), (0, -1, 0), (-1, 0, 0))
    # tank (Create.py:254-430)
    xbmin, xbmax = x_box - 0.5 * Nx_box * dr, x_box + 0.5 * Nx_box * dr
    ybmin, ybmax = -0.5 * Ny_box * dr, 0.5 * Ny_box * dr
    x0, y0, z0 = -(Nx - nx) * dr, -0.5 * Ny * dr, 0.0
    IX, IY, IZ = np.arange(Nx), np.arange(Ny), np.arange(Nz)
    q = grid(IX, IY).astype(np.float64)
    bx, by = x0 + (q[:, 0] + .5) * dr, y0 + (q[:, 1] + .5) * dr
    keep = ~((bx > xbmin) & (bx < xbmax) & (by > ybmin) & (by < ybmax))
    face(np.stack([bx[keep], by[keep], np.full(keep.sum(), z0)], 1), (0, 0, -1), (-1, 0, 0))
    face(np.stack([bx, by, np.full(len(q), z0 + Nz * dr)], 1), (0, 0, 1), (1, 0, 0))
    q = grid(IY, IZ).astype(np.float64)
    face(np.stack([np.full(len(q), x0), y0 + (q[:, 0] + .5) * dr, z0 + (q[:, 1] + .5) * dr], 1),
         (-1, 0, 0), (0, -1, 0))
    face(np.stack([np.full(len(q), x0 + Nx * dr), y0 + (q[:, 0] + .5) * dr,
                   z0 + (q[:, 1] + .5) * dr], 1), (1, 0, 0), (0, 1, 0))
    q = grid(IX, IZ).astype(np.float64)
    face(np.stack([x0 + (q[:, 0] + .5) * dr, np.full(len(q), y0), z0 + (q[:, 1] + .5) * dr], 1),
         (0, -1, 0), (1, 0, 0))
    face(np.stack([x0 + (q[:, 0] + .5) * dr, np.full(len(q), y0 + Ny * dr),
                   z0 + (q[:, 1] + .5) * dr], 1), (0, 1, 0), (-1, 0, 0))
    nset0 = sum(len(p[0]) for p in parts)
    # sensors (Create.py:432-470): set 1
    zz = Nz_box * dr
    spos = [(xbmax, 0.0, 0.021 + i * 0.04) for i in range(4)] + \
           [(xbmax - 0.021 - i * 0.04, 0.0, zz) for i in range(4)]
    snrm = [(-1, 0, 0)] * 4 + [(0, 0, 1)] * 4
    stan = [(0, 1, 0)] * 4 + [(1, 0, 0)] * 4
    parts.append((np.array(spos), np.array(snrm, float), np.array(stan, float),
                  np.zeros(8, np.int32)))

    pos = np.concatenate([p[0] for p in parts])
    N = len(pos)
    r = np.zeros((N, 4), np.float32); r[:, :3] = pos
    normal = np.zeros((N, 4), np.float32); normal[:, :3] = np.concatenate([p[1] for p in parts])
    tangent = np.zeros((N, 4), np.float32); tangent[:, :3] = np.concatenate([p[2] for p in parts])
    imove = np.concatenate([p[3] for p in parts]).astype(np.int32)
    rho = np.full(N, refd, np.float32); rho[:nfluid] = dens_f
    m = np.full(N, dr ** 2, np.float32); m[:nfluid] = dens_f * dr ** 3; m[nset0:] = 0.0
    iset = np.zeros(N, np.uint32); iset[nset0:] = 1
    s = 10.0 * hfac * dr
    dmin = np.array([-(L - l + s), -(0.5 * D + s), -s, 0.0], np.float32)
    dmax = np.array([l + s, 0.5 * D + s, H + s, 0.0], np.float32)
    hh = float(np.float32(np.float32(hfac) * np.float32(dr)))
    u = np.zeros((N, 4), np.float32)
    if seed is not None:  # optional perturbation so that every term is exercised
        rng = np.random.default_rng(seed)
        u[:nfluid, :3] = uscale * rng.uniform(-1, 1, (nfluid, 3))
        if jitter:  # off-lattice positions: halo / migration counts of a slab cut are arbitrary
            r[:nfluid, :3] += (jitter * dr * rng.uniform(-1, 1, (nfluid, 3))).astype(np.float32)
    return dict(
        dims=3, N=N, n_fluid=nfluid, h=hh, dr=float(np.float32(dr)), hfac=hfac, cs=cs, p0=0.0,
        support=2.0, refd=np.array([refd, refd], np.float32),
        visc_dyn=np.array([visc_dyn, visc_dyn], np.float32),
        delta=np.array([delta, delta], np.float32),
        g=np.array([0, 0, -g, 0], np.float32), domain_min=dmin, domain_max=dmax,
        courant=courant, dt_Ma=0.1, dt_min=float(np.float32(0.05 * courant * hh / cs)),
        id=np.arange(N, dtype=np.uint32), r=r, imove=imove, iset=iset, normal=normal,
        tangent=tangent, rho=rho, m=m, u=u, dudt=np.zeros((N, 4), np.float32),
        drhodt=np.zeros(N, np.float32),
    )


def spheric5_dam_break_2d(n=50000, hfac=3.0, seed=None):
    """BASELINE config 1: the 2-D SPHERIC test 5 dam break over a wet bed, geometry and field
    initialisation of examples/2D/spheric_testcase5_dambreak/src/Create.py:41-190: reservoir
    l0 x h0 = 0.38 x 0.15 m (n particles), wetted bed l1 x h1 = 9.55 x 0.038 m, floor and
    two walls of boundary-integral elements (imove = -3, m = dr), hydrostatic density."""
    g, cs, courant, refd = 9.81, 45.0, 0.1, 998.0
    delta, visc_dyn = 10.0, 0.000894
    h0, h1, l0, l1 = 15e-2, 38e-3, 38e-2, 955e-2
    H = 2.0 * h0
    dr = (l0 * h0 / n) ** 0.5

    def column(y_top):
        k = np.arange(int(np.ceil(y_top / dr)) + 2)
        y = 0.5 * dr + k * dr
        return y[y < y_top]

    # reservoir: x = -0.5 dr, -1.5 dr, ... > -l0 ; y = 0.5 dr ... < h0 (x outer, y inner)
    k = np.arange(int(np.ceil(l0 / dr)) + 2)
    xr = -0.5 * dr - k * dr
    xr = xr[xr > -l0]
    yr = column(h0)
    X, Y = np.meshgrid(xr, yr, indexing="ij")
    res = np.stack([X.ravel(), Y.ravel()], 1)
    xmin = xr[-1]
    k = np.arange(int(np.ceil(l1 / dr)) + 2)
    xb = 0.5 * dr + k * dr
    xb = xb[xb < l1]
    yb = column(h1)
    X, Y = np.meshgrid(xb, yb, indexing="ij")
    bed = np.stack([X.ravel(), Y.ravel()], 1)
    xmax = xb[-1]
    fluid = np.concatenate([res, bed])
    nf = len(fluid)
    # boundary elements: floor, left wall, right wall
    k = np.arange(int(np.ceil((l1 - xmin) / dr)) + 2)
    xf = xmin + k * dr
    xf = xf[xf < l1]
    floor = np.stack([xf, np.zeros_like(xf)], 1)
    yw = column(H)
    left = np.stack([np.full_like(yw, xmin - 0.5 * dr), yw], 1)
    right = np.stack([np.full_like(yw, xmax + 0.5 * dr), yw], 1)
    bnd = np.concatenate([floor, left, right])
    nrm = np.concatenate([np.tile([0.0, -1.0], (len(floor), 1)), np.tile([-1.0, 0.0], (len(left), 1)),
                          np.tile([1.0, 0.0], (len(right), 1))])
    N = nf + len(bnd)
    r = np.concatenate([fluid, bnd]).astype(np.float32)
    y = np.concatenate([fluid, bnd])[:, 1]
    press = refd * g * (h0 - y)
    press[nf + len(floor):] = np.maximum(0.0, press[nf + len(floor):])
    rho = (refd + press / cs ** 2).astype(np.float32)
    m = np.full(N, refd * dr ** 2, np.float32)
    m[nf:] = dr
    imove = np.ones(N, np.int32)
    imove[nf:] = -3
    normal = np.zeros((N, 2), np.float32)
    normal[nf:] = nrm
    u = np.zeros((N, 2), np.float32)
    if seed is not None:
        rng = np.random.default_rng(seed)
        u[:nf] = 0.05 * rng.uniform(-1, 1, (nf, 2))
    hh = float(np.float32(np.float32(hfac) * np.float32(dr)))
    return dict(
        dims=2, N=N, n_fluid=nf, h=hh, dr=float(np.float32(dr)), hfac=hfac, cs=cs, p0=0.0, support=2.0,
        refd=np.array([refd], np.float32), visc_dyn=np.array([visc_dyn], np.float32),
        delta=np.array([delta], np.float32), g=np.array([0, -g], np.float32),
        domain_min=np.array([-1.5 * l0, -0.5 * H], np.float32),
        domain_max=np.array([1.5 * l1, 2.0 * H], np.float32), courant=courant, dt_Ma=0.1,
        dt_min=float(np.float32(0.05 * courant * hh / cs)), id=np.arange(N, dtype=np.uint32), r=r,
        imove=imove, iset=np.zeros(N, np.uint32), normal=normal, tangent=np.zeros((N, 2), np.float32),
        rho=rho, m=m, u=u, dudt=np.zeros((N, 2), np.float32), drhodt=np.zeros(N, np.float32),
    )


def spheric9_tld_2d(n=10000, hfac=4.0, seed=None):
    """BASELINE config 4: the 2-D tuned liquid damper (SPHERIC test 9), geometry and field
    initialisation of examples/2D/spheric_testcase9_tld/src/Create.py:41-235: a tank L x H =
    0.9 x 0.508 m filled to h = 0.092 m (set 0: nx x ny fluid particles, hydrostatic density), closed
    by bottom, roof, left and right walls of boundary-integral elements (set 1: imove = -3, m = dr,
    outward normals), which the motion preset rotates around motion_r = (0, 0.47)."""
    g, cs, courant, refd = 9.81, 50.0, 0.1, 998.0
    delta, visc_dyn = 0.1, 0.000894
    H, L, hw = 0.508, 0.9, 0.092
    dr = (L * hw / n) ** 0.5
    nx, ny = int(round(L / dr)), int(round(hw / dr))
    nf = nx * ny
    h_fluid = ny * dr
    Nx, Ny = nx, int(round(H / dr)) + 1
    L, H = Nx * dr, Ny * dr
    k = np.arange(nf)
    fluid = np.stack([(k % nx) * dr - 0.5 * (L - dr), (k // nx) * dr + 0.5 * dr], 1)
    jx, jy = np.arange(Nx) + 0.5, np.arange(Ny) + 0.5
    bnd = np.concatenate([
        np.stack([jx * dr - 0.5 * L, np.zeros(Nx)], 1),            # bottom
        np.stack([jx * dr - 0.5 * L, np.full(Nx, Ny * dr)], 1),    # roof
        np.stack([np.full(Ny, -0.5 * L), jy * dr], 1),             # left
        np.stack([np.full(Ny, Nx * dr - 0.5 * L), jy * dr], 1)])   # right
    nrm = np.concatenate([np.tile([0.0, -1.0], (Nx, 1)), np.tile([0.0, 1.0], (Nx, 1)),
                          np.tile([-1.0, 0.0], (Ny, 1)), np.tile([1.0, 0.0], (Ny, 1))])
    nb = len(bnd)
    N = nf + nb
    y = np.concatenate([fluid, bnd])[:, 1]
    press = np.where(y <= h_fluid, refd * g * (h_fluid - y), 0.0)
    rho = (refd + press / cs ** 2).astype(np.float32)
    m = (rho.astype(np.float64) * dr ** 2).astype(np.float32)
    m[nf:] = dr
    imove = np.ones(N, np.int32)
    imove[nf:] = -3
    iset = np.zeros(N, np.uint32)
    iset[nf:] = 1
    normal = np.zeros((N, 2), np.float32)
    normal[nf:] = nrm
    u = np.zeros((N, 2), np.float32)
    if seed is not None:
        rng = np.random.default_rng(seed)
        u[:nf] = 0.05 * rng.uniform(-1, 1, (nf, 2))
    radius = (0.25 * L * L + H * H) ** 0.5
    hh = float(np.float32(np.float32(hfac) * np.float32(dr)))
    return dict(
        dims=2, N=N, n_fluid=nf, n_set0=nf, n_set1=nb, h=hh, dr=float(np.float32(dr)), hfac=hfac, cs=cs,
        p0=0.0, support=2.0, refd=np.array([refd, refd], np.float32),
        visc_dyn=np.array([visc_dyn, visc_dyn], np.float32), delta=np.array([delta, delta], np.float32),
        g=np.array([0, -g], np.float32), domain_min=np.array([-1.1 * radius, -0.5 * L], np.float32),
        domain_max=np.array([1.1 * radius, 1.1 * radius], np.float32), courant=courant, dt_Ma=0.1,
        dt_min=float(np.float32(0.05 * courant * hh / cs)), id=np.arange(N, dtype=np.uint32),
        r=np.concatenate([fluid, bnd]).astype(np.float32), imove=imove, iset=iset, normal=normal,
        tangent=np.zeros((N, 2), np.float32), rho=rho, m=m, u=u, dudt=np.zeros((N, 2), np.float32),
        drhodt=np.zeros(N, np.float32), motion_r=(0.0, 0.47),
    )


def spheric3_lid_driven_2d(nx=200, hfac=4.0, Re=1000.0):
    """The lid-driven cavity (SPHERIC test 3), geometry and fields of
    examples/2D/spheric_testcase3_liddriven/src/Create.py:41-230: a unit square of nx x nx fluid
    particles at rest (set 0, rho = refd = 1, background pressure p0 = 3 refd U^2), closed by four walls
    of nx boundary-integral elements each (set 1: imove = -3, m = dr, outward normals), the top one
    moving with u = (U, 0); no gravity; the no-slip set is noslip_iset = 1 (BINoSlip.xml)."""
    cs, courant, refd, L, U = 50.0, 0.1, 1.0, 1.0, 1.0
    dr = L / nx
    visc_dyn = refd * U * L / Re
    alpha = 8.0 * visc_dyn / (refd * hfac * dr * cs)
    delta = 1.0 if alpha < 0.03 else 0.0
    k = np.arange(nx)
    c = -0.5 * (L - dr) + k * dr
    X, Y = np.meshgrid(c, c, indexing="ij")                      # x outer, y inner
    fluid = np.stack([X.ravel(), Y.ravel()], 1)
    half = np.full(nx, 0.5 * L)
    bnd = np.concatenate([np.stack([c, half], 1), np.stack([c, -half], 1),       # top, bottom
                          np.stack([-half, c], 1), np.stack([half, c], 1)])      # left, right
    nrm = np.concatenate([np.tile([0.0, 1.0], (nx, 1)), np.tile([0.0, -1.0], (nx, 1)),
                          np.tile([-1.0, 0.0], (nx, 1)), np.tile([1.0, 0.0], (nx, 1))])
    nf, nb = len(fluid), len(bnd)
    N = nf + nb
    u = np.zeros((N, 2), np.float32)
    u[nf:nf + nx, 0] = U                                         # the lid
    m = np.full(N, refd * dr ** 2, np.float32)
    m[nf:] = dr
    imove = np.ones(N, np.int32)
    imove[nf:] = -3
    iset = np.zeros(N, np.uint32)
    iset[nf:] = 1
    normal = np.zeros((N, 2), np.float32)
    normal[nf:] = nrm
    hh = float(np.float32(np.float32(hfac) * np.float32(dr)))
    return dict(
        dims=2, N=N, n_fluid=nf, n_set0=nf, n_set1=nb, h=hh, dr=float(np.float32(dr)), hfac=hfac, cs=cs,
        p0=3.0 * refd * U ** 2, support=2.0, refd=np.array([refd, refd], np.float32),
        visc_dyn=np.array([visc_dyn, visc_dyn], np.float32), delta=np.array([delta, delta], np.float32),
        g=np.array([0, 0], np.float32), domain_min=np.array([-L, -L], np.float32),
        domain_max=np.array([L, L], np.float32), courant=courant, dt_Ma=0.1,
        dt_min=float(np.float32(0.05 * courant * hh / cs)), id=np.arange(N, dtype=np.uint32),
        r=np.concatenate([fluid, bnd]).astype(np.float32), imove=imove, iset=iset, normal=normal,
        tangent=np.zeros((N, 2), np.float32), rho=np.full(N, refd, np.float32), m=m, u=u,
        dudt=np.zeros((N, 2), np.float32), drhodt=np.zeros(N, np.float32),
    )


def souto2012_standing_wave_2d(ny=100, hfac=4.0, periods=8):
    """The viscous standing wave of Souto-Iglesias et al. 2012, geometry and fields of
    examples/2D/souto_etal_2012_standingwave/src/Create.py:40-238: a tank L x H = 2 x 1 of nx x ny =
    2 ny x ny fluid particles (x inner, y outer) moving with the linear standing-wave velocity field,
    hydrostatic density, a bottom of nx boundary-integral elements (imove = -3, m = dr, normal (0, -1)),
    and n + nx buffer particles (imove = -255, m = 0) parked beyond domain_max for the two symmetry
    planes x = 0 and x = L of templates/Symmetries.xml; one particles set."""
    g, cs, courant, refd, alpha, delta = 1.0, 50.0, 0.1, 1.0, 0.0, 10.0
    L = 2.0
    H = 0.5 * L
    k = 2.0 * math.pi / L
    omega = math.sqrt(g * k * math.tanh(k * H))
    eps, Re, sep = 0.1, 250.0, 2.0
    nx = 2 * ny
    dr = H / ny
    h = hfac * dr
    n = nx * ny
    visc_dyn = max(alpha / 8.0 * refd * h * cs, refd * H * math.sqrt(g * H) / Re)
    A = 0.5 * eps * H
    dmin = (-0.2 * L - 6.0 * sep * h, -0.2 * (H + A) - 6.0 * sep * h)
    dmax = (1.2 * L + 6.0 * sep * h, 1.2 * (H + A) + 6.0 * sep * h)
    nb = nx
    nbuf = n + nb
    N = n + nb + nbuf
    T = 2.0 * math.pi / omega
    nu = H * math.sqrt(g * H) / Re
    ekin0 = eps ** 2 * g * H ** 2 * L / 32 * 2
    epot0 = 0.5 * g * H * (L * H * refd)
    j = np.arange(n)
    px = (j % nx) * dr + 0.5 * dr
    py = (j // nx) * dr + 0.5 * dr
    y = py - H
    ku = eps * g * H * k / (2.0 * omega * math.cosh(k * H))
    r = np.zeros((N, 2), np.float64)
    u = np.zeros((N, 2), np.float64)
    rho = np.full(N, refd, np.float64)
    m = np.zeros(N, np.float64)
    imove = np.full(N, -255, np.int32)
    normal = np.zeros((N, 2), np.float32)
    r[:n, 0], r[:n, 1] = px, py
    u[:n, 0] = ku * np.sin(k * px) * np.cosh(k * (H + y))
    u[:n, 1] = -ku * np.cos(k * px) * np.sinh(k * (H + y))
    rho[:n] = refd + refd * g * (H - py) / cs ** 2
    m[:n] = rho[:n] * dr ** 2
    imove[:n] = 1
    r[n:n + nb, 0] = (np.arange(nb) + 0.5) * dr
    rho[n:n + nb] = refd + refd * g * H / cs ** 2
    m[n:n + nb] = dr
    imove[n:n + nb] = -3
    normal[n:n + nb] = (0.0, -1.0)
    r[n + nb:] = (dmax[0] + sep * h, dmax[1] + sep * h)
    hh = float(np.float32(np.float32(hfac) * np.float32(dr)))
    _ = nu
    return dict(
        dims=2, N=N, n_fluid=n, n_boundary=nb, n_buffer=nbuf, h=hh, dr=float(np.float32(dr)), hfac=hfac, cs=cs, p0=0.0,
        support=2.0, refd=np.array([refd], np.float32), visc_dyn=np.array([visc_dyn], np.float32),
        delta=np.array([delta], np.float32), g=np.array([0.0, -g], np.float32),
        domain_min=np.array(dmin, np.float32), domain_max=np.array(dmax, np.float32), courant=courant, dt_Ma=0.1,
        dt_min=float(np.float32(0.05 * courant * hh / cs)), t_end=periods * T,
        placeholders={"L": repr(L), "END_TIME": repr(periods * T), "E_KIN": repr(ekin0), "E_POT": repr(epot0)},
        id=np.arange(N, dtype=np.uint32), r=r.astype(np.float32), imove=imove, iset=np.zeros(N, np.uint32),
        normal=normal, tangent=np.zeros((N, 2), np.float32), rho=rho.astype(np.float32), m=m.astype(np.float32),
        u=u.astype(np.float32), dudt=np.zeros((N, 2), np.float32), drhodt=np.zeros(N, np.float32),
        L=L, period=T,
    )


def shock_point_2d(n=50000, hfac=2.0):
    """The circular blast of an ideal gas of examples/2D/shock_point/src/Create.py:39-152: a disc of radius
    R = 0.5 of particles on a square lattice of pitch dr = sqrt(pi R^2 / n) (x outer, y inner, both from -R),
    gas at rest with rho = 1.00001 everywhere and the pressure 2e5 inside R0 = 0.2, 1e5 outside (internal
    energy e = p / ((gamma - 1) rho), gamma = 1.4); one particles set, every particle fluid (imove = 1; the
    case-local bc.cl freezes the rim while the time scheme runs)."""
    courant, R, R0, gamma = 0.25, 0.5, 0.2, 1.4
    p1, p2, rho1, rho2 = 2.0e5, 1.0e5, 1.00001, 1.00001
    c1, c2 = math.sqrt(gamma * p1 / rho1), math.sqrt(gamma * p2 / rho2)
    cs = max(c1, c2)
    e1, e2 = p1 / ((gamma - 1.0) * rho1), p2 / ((gamma - 1.0) * rho2)
    dr = (math.pi * R ** 2 / n) ** 0.5
    h = hfac * dr
    xs = []
    x = -R
    while x < R:          # (Create.py:118-133: repeated addition, not an index times dr)
        xs.append(x)
        x += dr
    xs = np.array(xs, np.float64)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    rr = np.sqrt(X ** 2 + Y ** 2)
    keep = ~(rr > R)
    px, py, rad = X[keep], Y[keep], rr[keep]
    N = int(keep.sum())
    inner = rad < R0
    rho = np.where(inner, rho1, rho2)
    eint = np.where(inner, e1, e2)
    Rd = R + 4.0 * h
    hh = float(np.float32(h))
    return dict(
        dims=2, N=N, n_fluid=N, h=hh, dr=float(np.float32(dr)), hfac=hfac, cs=cs, p0=0.0, support=2.0,
        refd=np.array([0.0], np.float32), visc_dyn=np.array([0.0], np.float32), delta=np.array([0.0], np.float32),
        g=np.array([0.0, 0.0], np.float32), gamma=np.array([gamma], np.float32),
        domain_min=np.array((-Rd, -Rd), np.float32), domain_max=np.array((Rd, Rd), np.float32), courant=courant,
        t_end=R0 / cs, R=R, R0=R0, e1=e1, e2=e2, p1=p1, p2=p2,
        placeholders={"H": repr(h), "GAMMA": repr(gamma), "R": repr(R)},
        id=np.arange(N, dtype=np.uint32), r=np.stack([px, py], 1).astype(np.float32),
        imove=np.ones(N, np.int32), iset=np.zeros(N, np.uint32), normal=np.zeros((N, 2), np.float32),
        tangent=np.zeros((N, 2), np.float32), rho=rho.astype(np.float32), m=(rho * dr ** 2).astype(np.float32),
        u=np.zeros((N, 2), np.float32), dudt=np.zeros((N, 2), np.float32), drhodt=np.zeros(N, np.float32),
        eint=eint.astype(np.float32), deintdt=np.zeros(N, np.float32),
    )


def spheric2_dam_break_slab(n_total, hfac, rank, size, buffer_frac=0.1, boundary_margin=None, seed=None,
                            jitter=0.0, uscale=0.1):
    """BASELINE config 3 shape: the 3-D dam break cut in `size` slabs along y, the way
    examples/3D/spheric_testcase2_dambreak_mpi/src/Create.py:140-200 does it: rank k
    owns the fluid with y in (y_min + k dy, y_min + (k+1) dy), dy = (domain_max_y -
    domain_min_y) / size (templates/MPI.xml:24-38), carries boundary elements and the
    sensors itself (the reference replicates ALL of them on every rank; here only those
    within `boundary_margin` of the slab, default 4 h, which is every element a particle
    of the slab or of its halo can see) and ends set 0 with buffer particles
    (imove = -255 parked at domain_max, Create.py:170-190) that receive migrating
    particles.  The y limits of the domain hug the tank (2 dr of slack) so that the
    slabs hold the same amount of fluid.  Returns the case dict of this rank with
    n_set0 (fluid + boundary + buffer) and the global particle count.  `seed`: the perturbed
    state of spheric2_dam_break (the same on every rank: it is cut afterwards)."""
    c = spheric2_dam_break(n_total, hfac, seed=seed, jitter=jitter, uscale=uscale)
    dr, h = c["dr"], c["h"]
    dmin, dmax = c["domain_min"].copy(), c["domain_max"].copy()
    dmin[1], dmax[1] = -0.5 - 2.0 * dr, 0.5 + 2.0 * dr
    dy = (float(dmax[1]) - float(dmin[1])) / size
    y0, y1 = float(dmin[1]) + rank * dy, float(dmin[1]) + (rank + 1) * dy
    margin = 4.0 * h if boundary_margin is None else boundary_margin
    y = c["r"][:, 1]
    imove = c["imove"]
    fluid = np.flatnonzero((imove == 1) & (y > y0) & (y <= y1))
    bound = np.flatnonzero((imove == -3) & (y > y0 - margin) & (y <= y1 + margin))
    sens = np.flatnonzero(imove == 0)
    nbuf = max(1024, int(buffer_frac * len(fluid)))
    keep0 = np.concatenate([fluid, bound])
    N0 = len(keep0) + nbuf
    N = N0 + len(sens)
    out = dict(c)
    for k in ("r", "normal", "tangent", "u", "dudt"):
        a = np.zeros((N, 4), np.float32)
        a[:len(keep0)] = c[k][keep0]
        a[N0:] = c[k][sens]
        if k == "r":
            a[len(keep0):N0] = dmax
        out[k] = a
    for k, fill in (("rho", c["refd"][0]), ("drhodt", 0.0), ("m", c["refd"][0] * dr ** 3)):
        a = np.full(N, fill, np.float32)
        a[:len(keep0)] = c[k][keep0]
        a[N0:] = c[k][sens]
        out[k] = a
    im = np.full(N, -255, np.int32)
    im[:len(keep0)] = imove[keep0]
    im[N0:] = 0
    iset = np.zeros(N, np.uint32)
    iset[N0:] = 1
    out.update(imove=im, iset=iset, id=np.arange(N, dtype=np.uint32), N=N, n_set0=N0,
               n_fluid=len(fluid), n_fluid_global=int((imove == 1).sum()), n_buffer=nbuf,
               fluid_index=fluid, boundary_index=bound,
               domain_min=dmin, domain_max=dmax, slab=(y0, y1))
    return out


def lattice_slab(n_side, hfac, rank, size, buffer_frac=0.1, **kw):
    """This rank's share of `lattice(n_side, hfac)` for a run on `size` devices: the particles with
    z in (z0, z1], z slabs of equal thickness between domain_min_z = -2 and domain_max_z = n_side + 2
    (the planes of cases_xml/src/lattice_mpi_3d/Slabs.xml), followed by buffer particles
    (imove = -255 parked at domain_max, basic/setBuffer.xml:35-49) that give arrivals room.
    own = indices of the rows in the one-device lattice."""
    c = lattice(n_side, hfac, **kw)
    dmin, dmax = c["domain_min"].copy(), c["domain_max"].copy()
    dmin[2], dmax[2] = -2.0, n_side + 2.0
    dmin[3] = dmax[3] = 0.0
    dz = (float(dmax[2]) - float(dmin[2])) / size
    z0, z1 = float(dmin[2]) + rank * dz, float(dmin[2]) + (rank + 1) * dz
    z = c["r"][:, 2]
    own = np.flatnonzero((z > z0) & (z <= z1))
    nbuf = max(256, int(buffer_frac * len(own)))
    N = len(own) + nbuf
    out = dict(c)
    for k in ("r", "normal", "tangent", "u", "dudt"):
        a = np.zeros((N, 4), np.float32)
        a[:len(own)] = c[k][own]
        if k == "r":
            a[len(own):] = dmax
        out[k] = a
    for k, fill in (("rho", c["refd"][0]), ("drhodt", 0.0), ("m", c["m"][0])):
        a = np.full(N, fill, np.float32)
        a[:len(own)] = c[k][own]
        out[k] = a
    im = np.full(N, -255, np.int32)
    im[:len(own)] = 1
    out.update(imove=im, iset=np.zeros(N, np.uint32), id=np.arange(N, dtype=np.uint32), N=N, n_fluid=len(own),
               n_fluid_global=c["N"], n_buffer=nbuf, own=own, domain_min=dmin, domain_max=dmax, slab=(z0, z1))
    return out


def lattice_slab_local(n_xy, nz_local, hfac, rank, size, buffer_frac=0.02, seed=1234, cs=40.0):
    """Weak-scaled config 5 (8e6 particles per device): this rank's n_xy x n_xy x nz_local block of a
    lattice that is size * nz_local cells tall, generated locally (nothing of the other ranks'
    blocks is ever held: 6.4e7 particles on 8 devices), velocities from a generator seeded per rank.
    domain z = [0, size * nz_local] puts the planes of Slabs.xml exactly between the blocks.
    buffer_frac: rows kept free for arrivals (basic/setBuffer.xml); 2 % keeps the 200^3 block below
    2^23 rows, so that n_radix -- a power of two, the length of every mpi_* array and of the halo
    link-list -- is 8.4 M instead of 16.8 M (with 5 % every O(n_radix) tool of the halo path cost twice as much)."""
    rng = np.random.default_rng(seed + 7919 * rank)
    ax = [np.arange(n_xy, dtype=np.float32), np.arange(n_xy, dtype=np.float32),
          np.arange(nz_local, dtype=np.float32) + np.float32(rank * nz_local)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    n = g.shape[0]
    nbuf = max(256, int(buffer_frac * n))
    N = n + nbuf
    # x / y margins of four cells: the buffer rows and the unused rows of the halo list are parked at
    # domain_max (= r_max), all in ONE cell; a margin of less than two cells puts that cell in the
    # neighbourhood of the lattice's corner particles, whose sweeps then walk millions of parked
    # rows as candidates (hfac 3 with a margin of 10: 508 ms per step on the top rank)
    margin = max(10.0, 4.0 * 2.0 * float(hfac))
    dmin = np.array([-margin, -margin, 0.0, 0.0], np.float32)
    dmax = np.array([n_xy + margin, n_xy + margin, float(size * nz_local), 0.0], np.float32)
    r = np.zeros((N, 4), np.float32)
    r[:n, :3] = g + 0.5
    r[n:] = dmax
    u = np.zeros((N, 4), np.float32)
    u[:n, :3] = (0.01 * cs * rng.uniform(-1, 1, (n, 3))).astype(np.float32)
    refd = np.float32(1000.0)
    rho = np.full(N, refd, np.float32)
    rho[:n] = (refd * (1.0 + 1e-3 * rng.uniform(-1, 1, n))).astype(np.float32)
    imove = np.full(N, -255, np.int32)
    imove[:n] = 1
    z4 = np.zeros((N, 4), np.float32)
    return dict(
        dims=3, N=N, n_fluid=n, n_buffer=nbuf, h=float(hfac), dr=1.0, cs=float(cs), p0=0.0, support=2.0,
        refd=np.array([refd], np.float32), visc_dyn=np.array([1e-3], np.float32),
        delta=np.array([0.1], np.float32), g=np.zeros(4, np.float32), domain_min=dmin, domain_max=dmax,
        courant=0.25, dt_Ma=0.1, dt_min=1e-7, id=np.arange(N, dtype=np.uint32), r=r, imove=imove,
        iset=np.zeros(N, np.uint32), normal=z4, tangent=z4.copy(), rho=rho, m=np.full(N, refd, np.float32),
        u=u, dudt=z4.copy(), drhodt=np.zeros(N, np.float32),
        slab=(float(rank * nz_local), float((rank + 1) * nz_local)),
    )


MPI_PLANE_FIELDS = ("r", "u", "dudt", "rho", "drhodt", "m", "imove")


def mpi_plane(table, rank=None):
    """The reference's multi-device parity case (tests/2D/MPI_plane): `table` is its
    particles.dat (10000 x 10: r(2) u(2) dudt(2) rho drhodt m imove,
    main_serial.xml:26).  rank=None: the serial set; rank=0|1: that process' set as
    tests/2D/MPI_plane/cMake/Create.py:33-60 builds it -- its own particles (x <= 0
    for rank 0) followed by one buffer particle (imove = -255, parked at (1, 1)) per
    particle of the other process, so that both hold 10000 rows.
    Returns (arrays, own) with own = indices of the rows in the serial table."""
    t = np.asarray(table, np.float32)
    x = t[:, 0]
    if rank is None:
        own = np.arange(len(t))
        rows = t
    else:
        own = np.flatnonzero(x <= 0.0) if rank == 0 else np.flatnonzero(x > 0.0)
        buf = np.tile(np.array([1, 1, 0, 0, 0, 0, 1, 0, 0, -255], np.float32), (len(t) - len(own), 1))
        rows = np.concatenate([t[own], buf])
    arrays = dict(r=np.ascontiguousarray(rows[:, 0:2]), u=np.ascontiguousarray(rows[:, 2:4]),
                  dudt=np.ascontiguousarray(rows[:, 4:6]), rho=np.ascontiguousarray(rows[:, 6]),
                  drhodt=np.ascontiguousarray(rows[:, 7]), m=np.ascontiguousarray(rows[:, 8]),
                  imove=np.ascontiguousarray(rows[:, 9].astype(np.int32)))
    return arrays, own
