"""Pure-numpy scene builders producing the POD arrays of include/playrho_b200.h directly (no
reference library needed; used by smoke(), bench.py and the large-scale property tests).
Shapes follow SURVEY.md §8(d). Mass data restates MassData.cpp (disk: pi r^2 rho, I = m(r^2/2 +
|c|^2); polygon: triangle-fan area/centroid/inertia) in float64 then rounds to float32 -- parity
tests never depend on these values because GPU, oracle and reference all consume the same PODs."""
import numpy as np
from playrho_b200 import abi

GRAVITY = (0.0, -9.8)
DEFAULT_VERTEX_RADIUS = 0.01  # 2 * linearSlop, PolygonShapeConf.hpp:68 / EdgeShapeConf.hpp:60


class SceneBuilder:
    def __init__(self):
        self.shapes, self.children, self.verts, self.normals = [], [], [], []
        self.bodies, self.body_shapes = [], []
        self.mass = []  # per shape: (mass, I about shape origin, center)

    # ---- shapes
    def _add_shape(self, child_list, friction, restitution, density, sensor=0):
        sid = len(self.shapes)
        self.shapes.append((friction, restitution, density, 1, 0xFFFF, 0, sensor, len(self.children), len(child_list)))
        for (vs, ns, r) in child_list:
            self.children.append((len(self.verts), len(vs), r))
            self.verts.extend(vs)
            self.normals.extend(ns)
        return sid

    def disk(self, radius, friction=0.2, restitution=0.0, density=0.0):
        sid = self._add_shape([([(0.0, 0.0)], [(0.0, 0.0)], radius)], friction, restitution, density)
        m = density * np.pi * radius * radius
        self.mass.append((m, m * (radius * radius / 2.0), (0.0, 0.0)))
        return sid

    @staticmethod
    def _fwd_normals(vs):
        vs = np.asarray(vs, dtype=np.float32)
        nxt = np.roll(vs, -1, axis=0)
        e = (nxt - vs).astype(np.float32)
        n = np.stack([e[:, 1], -e[:, 0]], 1).astype(np.float32)
        ms = (n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]).astype(np.float32)
        inv = (np.float32(1.0) / np.sqrt(ms).astype(np.float32)).astype(np.float32)
        return (n * inv[:, None]).astype(np.float32)

    def edge(self, p0, p1, friction=0.2, restitution=0.0, vertex_radius=DEFAULT_VERTEX_RADIUS):
        vs = [tuple(map(float, p0)), tuple(map(float, p1))]
        ns = self._fwd_normals(vs)
        sid = self._add_shape([(vs, [tuple(x) for x in ns], vertex_radius)], friction, restitution, 0.0)
        self.mass.append((0.0, 0.0, (0.0, 0.0)))
        return sid

    def chain(self, pts, friction=0.2, restitution=0.0, vertex_radius=DEFAULT_VERTEX_RADIUS):
        kids = []
        for a, b in zip(pts[:-1], pts[1:]):
            vs = [tuple(map(float, a)), tuple(map(float, b))]
            ns = self._fwd_normals(vs)
            kids.append((vs, [tuple(x) for x in ns], vertex_radius))
        sid = self._add_shape(kids, friction, restitution, 0.0)
        self.mass.append((0.0, 0.0, (0.0, 0.0)))
        return sid

    def polygon(self, pts, friction=0.2, restitution=0.0, density=0.0, vertex_radius=DEFAULT_VERTEX_RADIUS):
        """pts: convex, counter-clockwise."""
        vs = [tuple(map(float, p)) for p in pts]
        ns = self._fwd_normals(vs)
        sid = self._add_shape([(vs, [tuple(x) for x in ns], vertex_radius)], friction, restitution, density)
        p = np.asarray(vs, dtype=np.float64)
        q = np.roll(p, -1, axis=0)
        cr = p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0]
        area = cr.sum() / 2.0
        c = ((p + q) * cr[:, None]).sum(0) / (6.0 * area) if area > 0 else p.mean(0)
        inert = (cr * ((p * p).sum(1) + (p * q).sum(1) + (q * q).sum(1))).sum() / 12.0
        m = density * area
        self.mass.append((m, density * inert, (float(c[0]), float(c[1]))))
        return sid

    def box(self, hx, hy, **kw):
        # SetAsBox: CCW from bottom-right (PolygonShapeConf.cpp:45-56)
        return self.polygon([(hx, -hy), (hx, hy), (-hx, hy), (-hx, -hy)], **kw)

    # ---- bodies
    def body(self, btype, pos=(0.0, 0.0), angle=0.0, vel=(0.0, 0.0), w=0.0, acc=(0.0, 0.0), shapes=(),
             allow_sleep=True, bullet=False):
        m = sum(self.mass[s][0] for s in shapes)
        inv_m, inv_i, lc = 0.0, 0.0, (0.0, 0.0)
        if btype == 2:
            if m > 0:
                cx = sum(self.mass[s][0] * self.mass[s][2][0] for s in shapes) / m
                cy = sum(self.mass[s][0] * self.mass[s][2][1] for s in shapes) / m
                inert = sum(self.mass[s][1] for s in shapes) - m * (cx * cx + cy * cy)
                inv_m, lc = 1.0 / m, (cx, cy)
                inv_i = 1.0 / inert if inert > 0 else 0.0
            else:
                inv_m = 1.0  # zero-density dynamic body gets 1 kg (WorldBody.cpp:376)
        c, s = np.float32(np.cos(np.float32(angle))), np.float32(np.sin(np.float32(angle)))
        wx = np.float32(pos[0]) + (c * np.float32(lc[0]) - s * np.float32(lc[1]))
        wy = np.float32(pos[1]) + (s * np.float32(lc[0]) + c * np.float32(lc[1]))
        flags = abi.BODY_ENABLED
        if btype == 2:
            flags |= abi.BODY_VELOCITY | abi.BODY_ACCELERATION | abi.BODY_AWAKE
        elif btype == 1:
            flags |= abi.BODY_VELOCITY | abi.BODY_IMPENETRABLE | abi.BODY_AWAKE
        else:
            flags |= abi.BODY_IMPENETRABLE
        if allow_sleep:
            flags |= abi.BODY_AUTOSLEEP
        if bullet:
            flags |= abi.BODY_IMPENETRABLE
        self.bodies.append((pos[0], pos[1], c, s, wx, wy, angle, wx, wy, angle, lc[0], lc[1], 0.0, vel[0], vel[1], w,
                            acc[0] if btype == 2 else 0.0, acc[1] if btype == 2 else 0.0, 0.0, inv_m, inv_i, 0.0, 0.0, 0.0, flags))
        self.body_shapes.append(list(shapes))
        return len(self.bodies) - 1

    def finish(self):
        shapes = np.array(self.shapes, dtype=abi.SHAPE) if self.shapes else np.zeros(0, abi.SHAPE)
        children = np.array(self.children, dtype=abi.CHILD)
        verts = np.array(self.verts, dtype=np.float32).reshape(-1, 2).copy().view(abi.VEC2).reshape(-1)
        normals = np.array(self.normals, dtype=np.float32).reshape(-1, 2).copy().view(abi.VEC2).reshape(-1)
        bodies = np.array(self.bodies, dtype=abi.BODY)
        offs = np.zeros(len(self.bodies) + 1, np.uint32)
        offs[1:] = np.cumsum([len(s) for s in self.body_shapes])
        ids = np.array([s for b in self.body_shapes for s in b], dtype=np.uint32)
        if len(ids) == 0:
            ids = np.zeros(1, np.uint32)
        return dict(shapes=shapes, children=children, verts=verts, normals=normals, bodies=bodies,
                    shape_offsets=offs, shape_ids=ids)


def disks_box(n, seed=12345, radius=0.5):
    """Config 3 recipe: 4 edges on one static body forming W x 2W, n disks on a jittered lattice."""
    rng = np.random.RandomState(seed)
    b = SceneBuilder()
    side = int(np.ceil(np.sqrt(n)))
    W = 1.1 * side
    edges = [b.edge((0.0, 0.0), (W, 0.0)), b.edge((W, 0.0), (W, 2 * W)), b.edge((W, 2 * W), (0.0, 2 * W)),
             b.edge((0.0, 2 * W), (0.0, 0.0))]
    b.body(0, shapes=edges)
    disk = b.disk(radius, density=1.0)
    jit = rng.uniform(-0.02, 0.02, (n, 2))
    for k in range(n):
        i, j = k % side, k // side
        b.body(2, pos=(0.6 + i * 1.05 + jit[k, 0], 0.6 + j * 1.05 + jit[k, 1]), acc=GRAVITY, shapes=[disk])
    return b.finish()


def pyramid(rows):
    """Config 2 recipe (Testbed/Tests/Pyramid.cpp): edge ground + rows of 1 m boxes, density 5."""
    b = SceneBuilder()
    g = b.edge((-60.0, 0.0), (60.0, 0.0))
    b.body(0, shapes=[g])
    box = b.box(0.5, 0.5, density=5.0)
    x = np.array([-7.0, 0.75], dtype=np.float32)
    dx = np.array([0.5625, 1.25], dtype=np.float32)
    dy = np.array([1.125, 0.0], dtype=np.float32)
    for i in range(rows):
        y = x.copy()
        for _ in range(i, rows):
            b.body(2, pos=(float(y[0]), float(y[1])), acc=GRAVITY, shapes=[box])
            y = (y + dy).astype(np.float32)
        x = (x + dx).astype(np.float32)
    return b.finish()


def mixed(n, seed=12345, tile_segments=256, tile_width=64.0):
    """Config 4 recipe: tiled static bodies each one sawtooth chain of `tile_segments`
    segments; n dynamic bodies, 50% disks r~U(0.2,0.5), 50% convex k-gons (k~U{4..8}) on a circle."""
    rng = np.random.RandomState(seed)
    b = SceneBuilder()
    cols = int(np.ceil(np.sqrt(n * 4.0)))      # wide strip: 4:1 aspect
    rows = int(np.ceil(n / cols))
    width = cols * 1.1 + 2.0
    ntiles = int(np.ceil(width / tile_width))
    for t in range(ntiles):
        xs = np.linspace(t * tile_width - 1.0, (t + 1) * tile_width - 1.0, tile_segments + 1)
        ys = np.where(np.arange(tile_segments + 1) % 2 == 0, 0.0, 0.5)
        c = b.chain(list(zip(xs, ys)))
        b.body(0, shapes=[c])
    wall_l = b.edge((-1.0, 0.0), (-1.0, rows * 1.1 + 8.0))
    wall_r = b.edge((ntiles * tile_width - 1.0, 0.0), (ntiles * tile_width - 1.0, rows * 1.1 + 8.0))
    b.body(0, shapes=[wall_l, wall_r])
    kinds = rng.rand(n) < 0.5
    radii = rng.uniform(0.2, 0.5, n)
    jit = rng.uniform(-0.02, 0.02, (n, 2))
    angs = rng.uniform(-3.0, 3.0, n)
    for k in range(n):
        i, j = k % cols, k // cols
        pos = (0.5 + i * 1.1 + jit[k, 0], 2.0 + j * 1.1 + jit[k, 1])
        if kinds[k]:
            s = b.disk(radii[k], density=1.0)
        else:
            kk = rng.randint(4, 9)
            a = np.sort(rng.uniform(0, 2 * np.pi, kk))
            # keep it convex and non-degenerate: spread angles at least 0.3 rad apart
            a = np.linspace(0, 2 * np.pi, kk, endpoint=False) + 0.6 * (a - np.linspace(0, 2 * np.pi, kk, endpoint=False)) / kk
            s = b.polygon(np.stack([radii[k] * np.cos(a), radii[k] * np.sin(a)], 1), density=1.0)
        b.body(2, pos=pos, angle=angs[k], acc=GRAVITY, shapes=[s])
    return b.finish()


def mixed_fast(n, seed=12345, tile_segments=256, tile_width=64.0, drop_height=2.0, strip_width=None):
    """Vectorised version of mixed() for large n (config 4: 1M bodies): same recipe, one shape per
    dynamic body, built without per-body python loops. strip_width (m): width of the terrain strip the lattice is dropped
    over (4 000 m for 1M bodies); None = a 4:1 lattice."""
    rng = np.random.RandomState(seed)
    f32 = np.float32
    cols = int(np.ceil(np.sqrt(n * 4.0))) if strip_width is None else max(1, int((strip_width - 2.0) / 1.1))
    rows = int(np.ceil(n / cols))
    width = cols * 1.1 + 2.0
    ntiles = int(np.ceil(width / tile_width))
    # --- static part through the scalar builder (few thousand segments)
    b = SceneBuilder()
    for t in range(ntiles):
        xs = np.linspace(t * tile_width - 1.0, (t + 1) * tile_width - 1.0, tile_segments + 1)
        ys = np.where(np.arange(tile_segments + 1) % 2 == 0, 0.0, 0.5)
        b.body(0, shapes=[b.chain(list(zip(xs, ys)))])
    wl = b.edge((-1.0, 0.0), (-1.0, rows * 1.1 + 8.0))
    wr = b.edge((ntiles * tile_width - 1.0, 0.0), (ntiles * tile_width - 1.0, rows * 1.1 + 8.0))
    b.body(0, shapes=[wl, wr])
    st = b.finish()
    ns0, nc0, nv0, nb0 = len(st["shapes"]), len(st["children"]), len(st["verts"]), len(st["bodies"])
    # --- dynamic part, vectorised
    is_disk = rng.rand(n) < 0.5
    radii = rng.uniform(0.2, 0.5, n)
    jit = rng.uniform(-0.02, 0.02, (n, 2))
    angs = rng.uniform(-3.0, 3.0, n).astype(f32)
    kk = np.where(is_disk, 1, rng.randint(4, 9, n)).astype(np.int64)
    voff = np.concatenate([[0], np.cumsum(kk)[:-1]]).astype(np.int64)
    nv = int(kk.sum())
    vx = np.zeros(nv, f32); vy = np.zeros(nv, f32)
    base = rng.uniform(0, 1, (n, 8))
    area = np.zeros(n); cx = np.zeros(n); cy = np.zeros(n); inert = np.zeros(n)
    P = np.zeros((n, 8, 2))
    for j in range(8):
        m = (~is_disk) & (kk > j)
        a = (2 * np.pi * (j + 0.2 + 0.6 * base[:, j]) / np.maximum(kk, 1))
        P[:, j, 0] = np.where(m, radii * np.cos(a), 0.0); P[:, j, 1] = np.where(m, radii * np.sin(a), 0.0)
        vx[(voff + j)[m]] = P[m, j, 0].astype(f32); vy[(voff + j)[m]] = P[m, j, 1].astype(f32)
    Pf = P.astype(f32).astype(np.float64)
    for j in range(8):
        m = (~is_disk) & (kk > j)
        jn = np.where(j + 1 < kk, j + 1, 0)
        p = Pf[:, j, :]; q = Pf[np.arange(n), jn, :]
        cr = np.where(m, p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0], 0.0)
        area += cr / 2.0
        cx += (p[:, 0] + q[:, 0]) * cr; cy += (p[:, 1] + q[:, 1]) * cr
        inert += cr * ((p * p).sum(1) + (p * q).sum(1) + (q * q).sum(1)) / 12.0
    poly = ~is_disk
    cx = np.where(poly, cx / np.where(poly, 6.0 * area, 1.0), 0.0); cy = np.where(poly, cy / np.where(poly, 6.0 * area, 1.0), 0.0)
    mass = np.where(is_disk, np.pi * radii * radii, area)
    I0 = np.where(is_disk, mass * radii * radii / 2.0, inert)
    Ic = I0 - mass * (cx * cx + cy * cy)
    # normals (forward perpendicular of each edge, normalised in float32)
    nx = np.zeros(nv, f32); ny = np.zeros(nv, f32)
    owner = np.repeat(np.arange(n), kk)
    local = np.arange(nv) - voff[owner]
    nxt = voff[owner] + np.where(local + 1 < kk[owner], local + 1, 0)
    ex = (vx[nxt] - vx).astype(f32); ey = (vy[nxt] - vy).astype(f32)
    pm = ~is_disk[owner]
    fx, fy = ey, (-ex).astype(f32)
    ms = (fx * fx + fy * fy).astype(f32)
    inv = np.where(pm, (f32(1.0) / np.sqrt(np.where(pm, ms, 1)).astype(f32)), 0).astype(f32)
    nx = (fx * inv).astype(f32); ny = (fy * inv).astype(f32)
    shapes = np.zeros(n, abi.SHAPE)
    shapes["friction"] = 0.2; shapes["density"] = 1.0; shapes["filter_category"] = 1; shapes["filter_mask"] = 0xFFFF
    shapes["child_offset"] = nc0 + np.arange(n); shapes["child_count"] = 1
    children = np.zeros(n, abi.CHILD)
    children["vertex_offset"] = nv0 + voff; children["vertex_count"] = kk
    children["vertex_radius"] = np.where(is_disk, radii, DEFAULT_VERTEX_RADIUS).astype(f32)
    verts = np.zeros(nv, abi.VEC2); verts["x"] = vx; verts["y"] = vy
    normals = np.zeros(nv, abi.VEC2); normals["x"] = nx; normals["y"] = ny
    k = np.arange(n)
    px = (0.5 + (k % cols) * 1.1 + jit[:, 0]).astype(f32); py = (drop_height + (k // cols) * 1.1 + jit[:, 1]).astype(f32)
    c = np.cos(angs).astype(f32); s = np.sin(angs).astype(f32)
    lcx = cx.astype(f32); lcy = cy.astype(f32)
    wx = (px + (c * lcx - s * lcy)).astype(f32); wy = (py + (s * lcx + c * lcy)).astype(f32)
    bodies = np.zeros(n, abi.BODY)
    bodies["xf_px"] = px; bodies["xf_py"] = py; bodies["xf_qc"] = c; bodies["xf_qs"] = s
    for f, v in (("pos0_x", wx), ("pos0_y", wy), ("pos0_a", angs), ("pos1_x", wx), ("pos1_y", wy), ("pos1_a", angs),
                 ("lc_x", lcx), ("lc_y", lcy)):
        bodies[f] = v
    bodies["acc_x"] = GRAVITY[0]; bodies["acc_y"] = GRAVITY[1]
    bodies["inv_mass"] = (1.0 / mass).astype(f32); bodies["inv_rot_i"] = (1.0 / Ic).astype(f32)
    bodies["flags"] = abi.BODY_ENABLED | abi.BODY_VELOCITY | abi.BODY_ACCELERATION | abi.BODY_AWAKE | abi.BODY_AUTOSLEEP
    offs = np.concatenate([st["shape_offsets"], st["shape_offsets"][-1] + 1 + np.arange(n)]).astype(np.uint32)
    ids_static = st["shape_ids"][:st["shape_offsets"][-1]]
    ids = np.concatenate([ids_static, ns0 + np.arange(n)]).astype(np.uint32)
    return dict(shapes=np.concatenate([st["shapes"], shapes]), children=np.concatenate([st["children"], children]),
                verts=np.concatenate([st["verts"], verts]), normals=np.concatenate([st["normals"], normals]),
                bodies=np.concatenate([st["bodies"], bodies]), shape_offsets=offs, shape_ids=ids)


def tumbler(n_squares=0, seed=12345, spread=4.0):
    """Config 5 unit (Benchmark/BenchmarkMain.cpp:2614-2722, TumblerPlayRho): static anchor body, dynamic enclosure of
    four 1 x 20 m boxes (density 5, sleeping not allowed) at (0,10) turned by a motorised revolute joint (0.05 pi rad/s,
    maxMotorTorque 1e5), plus n_squares 0.25 m squares (density 1). The benchmark drops its squares one per step at
    (0,10) (see tumbler_square_row); n_squares > 0 pre-places jittered squares for the tests."""
    rng = np.random.RandomState(seed)
    b = SceneBuilder()
    g = b.body(0)

    def wall(hx, hy, cx, cy):
        return b.polygon([(cx + hx, cy - hy), (cx + hx, cy + hy), (cx - hx, cy + hy), (cx - hx, cy - hy)], density=5.0)
    walls = [wall(0.5, 10.0, 10.0, 0.0), wall(0.5, 10.0, -10.0, 0.0), wall(10.0, 0.5, 0.0, 10.0), wall(10.0, 0.5, 0.0, -10.0)]
    e = b.body(2, pos=(0.0, 10.0), allow_sleep=False, shapes=walls)  # (BodyConf default: no linear acceleration)
    sq = b.box(0.125, 0.125, density=1.0)
    for _ in range(n_squares):
        b.body(2, pos=(rng.uniform(-spread, spread), 10.0 + rng.uniform(-spread, spread)), acc=GRAVITY, shapes=[sq])
    sc = b.finish()
    j = np.zeros(1, abi.REVOLUTE)
    j["body_a"], j["body_b"] = g, e
    j["local_anchor_a"]["x"], j["local_anchor_a"]["y"] = 0.0, 10.0
    j["enable_motor"], j["motor_speed"], j["max_motor_torque"] = 1, np.float32(0.05 * np.pi), 1e5
    sc["joints"] = j
    sc["square_shape"] = sq
    return sc


def tumbler_step_conf(conf):
    """TumblerPlayRho's StepConf (BenchmarkMain.cpp:2636-2657): the defaults except these two."""
    conf["displace_multiplier"] = 4.0
    conf["max_rotation"] = np.float32(0.5 * np.pi)
    return conf


def tumbler_square_row(sc):
    """One AddSquare() body row (BenchmarkMain.cpp:2716-2722): a square at (0,10) with earthly gravity."""
    b = SceneBuilder()
    b.mass = [None] * sc["square_shape"] + [(0.0625 * 1.0, 0.0625 * (0.0625 + 0.0625) / 12.0, (0.0, 0.0))]
    b.body(2, pos=(0.0, 10.0), acc=GRAVITY, shapes=[sc["square_shape"]])
    return np.array(b.bodies, dtype=abi.BODY)


def merge_worlds(scenes, layer_filters=False):
    """Concatenate independent scenes into one arena scene (world-major body order). Returns (scene, world index of
    each body). layer_filters=True gives world w's shapes the collision Filter {category = mask = 1 << w}: in a plain
    (reference / oracle) world that is what keeps the worlds apart; the batch arena does not need it."""
    out = dict(shapes=[], children=[], verts=[], normals=[], bodies=[], shape_ids=[], joints=[])
    gears = []  # the whole values of gear joints, by joint id of the merged table (abi.GEAR; rows of other types stay zeroed)
    offs, world = [0], []
    ns = nc = nv = nb = 0
    for wi, sc in enumerate(scenes):
        sh = sc["shapes"].copy()
        sh["child_offset"] += nc
        if layer_filters:
            assert wi < 16
            sh["filter_category"] = 1 << wi
            sh["filter_mask"] = 1 << wi
        ch = sc["children"].copy()
        ch["vertex_offset"] += nv
        out["shapes"].append(sh); out["children"].append(ch)
        out["verts"].append(sc["verts"]); out["normals"].append(sc["normals"])
        out["bodies"].append(sc["bodies"])
        n_ids = int(sc["shape_offsets"][-1])
        out["shape_ids"].append(sc["shape_ids"][:n_ids] + ns)
        offs.extend((sc["shape_offsets"][1:].astype(np.int64) + offs[-1]).tolist())
        if "joints" in sc and len(sc["joints"]):
            j = sc["joints"].copy()
            j["body_a"] += nb; j["body_b"] += nb
            out["joints"].append(j)
            g = np.zeros(len(j), abi.GEAR)
            if "gears" in sc and len(sc["gears"]) == len(j):
                g = np.ascontiguousarray(sc["gears"], dtype=abi.GEAR).copy()
                live = g["body_a"] != 0xFFFFFFFF
                for f in ("body_a", "body_b", "body_c", "body_d"):
                    g[f][live] += nb
            gears.append(g)
        world.extend([wi] * len(sc["bodies"]))
        ns += len(sc["shapes"]); nc += len(sc["children"]); nv += len(sc["verts"]); nb += len(sc["bodies"])
    merged = {k: np.concatenate(v) if len(v) else np.zeros(0, abi.REVOLUTE) for k, v in out.items()}
    merged["gears"] = np.concatenate(gears) if gears else np.zeros(0, abi.GEAR)
    merged["shape_offsets"] = np.array(offs, dtype=np.uint32)
    merged["shape_ids"] = merged["shape_ids"].astype(np.uint32)
    if len(merged["shape_ids"]) == 0:
        merged["shape_ids"] = np.zeros(1, np.uint32)
    return merged, np.array(world, dtype=np.uint32)
