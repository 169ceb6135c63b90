"""unomol_b200/basis.py -- patin.dat reader and synthetic-cluster generator (host side, numpy only).

Mirrors the reference's input format and normalisation so the same files drive both programs:
  * file layout                      reference Basis.hpp:181-255
  * contracted-shell normalisation   reference Basis.hpp:56-75 (Shell::normalize)
  * basis-function offsets           reference Basis.hpp:226-232
  * eps floor DBL_EPSILON*norb^2/2   reference Basis.hpp:249-254
The flattened arrays are exactly what include/unomol_b200.h's unomol_basis_desc carries.
"""
import os
import numpy as np


def ncart(l):
    return (l + 1) * (l + 2) // 2


class Basis:
    def __init__(self):
        self.nshell = self.nbf = self.ncen = self.maxl = self.nelec = self.maxits = 0
        self.eps = 1e-10
        self.int_flag = [0, 0]; self.scf_flag = [2, 1, 0]; self.prt_flag = [0, 0, 0]
        self.charge = np.zeros(0); self.xyz = np.zeros((0, 3))
        self.npr = np.zeros(0, np.int32); self.lv = np.zeros(0, np.int32); self.cen = np.zeros(0, np.int32)
        self.off = np.zeros(0, np.int32); self.poff = np.zeros(0, np.int32)
        self.alpha = np.zeros(0); self.coef_raw = np.zeros(0); self.coef = np.zeros(0)

    @property
    def no2(self):
        return self.nbf * (self.nbf + 1) // 2

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_patin(cls, path):
        tok = open(path).read().split()
        it = iter(tok)
        nxt = lambda: next(it)
        b = cls()
        b.nshell, b.nbf, b.ncen, b.maxl = int(nxt()), int(nxt()), int(nxt()), int(nxt())
        b.nelec, b.maxits = int(nxt()), int(nxt())
        b.eps = float(nxt())
        b.int_flag = [int(nxt()), int(nxt())]
        b.scf_flag = [int(nxt()), int(nxt()), int(nxt())]
        b.prt_flag = [int(nxt()), int(nxt()), int(nxt())]
        if b.maxl > 4:
            raise ValueError("Angular Momentum is too large for present program")
        b.charge = np.zeros(b.ncen); b.xyz = np.zeros((b.ncen, 3))
        for i in range(b.ncen):
            b.charge[i] = float(nxt())
            b.xyz[i] = [float(nxt()), float(nxt()), float(nxt())]
        shells = []
        for _ in range(b.nshell):
            npr, l, cen = int(nxt()), int(nxt()), int(nxt())
            al = np.zeros(npr); co = np.zeros(npr)
            for k in range(npr):
                al[k] = float(nxt()); co[k] = float(nxt())
            shells.append((l, cen, al, co))
        b._set_shells(shells)
        return b

    def _set_shells(self, shells):
        self.nshell = len(shells)
        self.npr = np.array([len(s[2]) for s in shells], np.int32)
        self.lv = np.array([s[0] for s in shells], np.int32)
        self.cen = np.array([s[1] for s in shells], np.int32)
        self.off = np.zeros(self.nshell, np.int32); self.poff = np.zeros(self.nshell, np.int32)
        o = p = 0
        for i, s in enumerate(shells):
            self.off[i] = o; self.poff[i] = p
            o += ncart(s[0]); p += len(s[2])
        self.nbf = o
        self.maxl = max(int(self.lv.max()), self.maxl)
        self.alpha = np.concatenate([s[2] for s in shells])
        self.coef_raw = np.concatenate([s[3] for s in shells])
        self.coef = np.zeros_like(self.coef_raw)
        for i in range(self.nshell):
            sl = slice(self.poff[i], self.poff[i] + self.npr[i])
            self.coef[sl] = normalize_shell(int(self.lv[i]), self.alpha[sl], self.coef_raw[sl])
        xeps = np.finfo(float).eps * self.nbf * self.nbf * 0.5
        if self.eps < xeps:
            self.eps = xeps

    def write_patin(self, path):
        """patin.dat in the reference's layout (so the reference binary can read our synthetic clusters)"""
        with open(path, "w") as f:
            f.write("\n%7d%7d%7d%3d\n" % (self.nshell, self.nbf, self.ncen, self.maxl))
            f.write("%7d%7d\n" % (self.nelec, self.maxits))
            f.write("%24.16e\n" % self.eps_in if hasattr(self, "eps_in") else "%24.16e\n" % self.eps)
            f.write(" %d %d\n %d %d %d\n %d %d %d\n" % tuple(self.int_flag + self.scf_flag + self.prt_flag))
            for i in range(self.ncen):
                f.write("%15.10f%15.10f%15.10f%15.10f\n" % (self.charge[i], *self.xyz[i]))
            for i in range(self.nshell):
                f.write("%3d%3d%5d\n" % (self.npr[i], self.lv[i], self.cen[i]))
                for k in range(self.npr[i]):
                    f.write("%24.10f%24.10f\n" % (self.alpha[self.poff[i] + k], self.coef_raw[self.poff[i] + k]))

    def nuclear_repulsion(self):
        """reference RHF.hpp:273-290"""
        e = 0.0
        for i in range(self.ncen):
            for j in range(i + 1, self.ncen):
                e += self.charge[i] * self.charge[j] / np.linalg.norm(self.xyz[i] - self.xyz[j])
        return e


def normalize_shell(l, al, co):
    """Shell::normalize, reference Basis.hpp:56-75 (the double sum uses the raw coefficients)."""
    twofact, piterm = 2.8284271247461903, 5.568327996831707
    lpow = 1.5 + l
    a1, a2 = np.meshgrid(al, al, indexing="ij")
    c1, c2 = np.meshgrid(co, co, indexing="ij")
    s = float(np.sum(c1 * c2 * (np.sqrt(a1 * a2) / (a1 + a2)) ** lpow)) * twofact
    s = 1.0 / np.sqrt(s)
    return co * s * np.sqrt((2 * al) ** lpow / piterm)


# ---------------------------------------------------------------------- synthetic water clusters
# 6-31G exponents / contraction coefficients for O and H as shipped in the reference's basis library
# (pbas.lib:2130-2150 for O, pbas.lib:24-31 for H); SURVEY.md 8(d) defines the cluster geometry.
O_631G = [
    (0, [5484.6716600, 825.2349460, 188.0469580, 52.9645000, 16.8975704, 5.7996353],
        [0.0018311, 0.0139502, 0.0684451, 0.2327143, 0.4701929, 0.3585209]),
    (0, [15.5396162, 3.5999336, 1.0137618], [-0.1107775, -0.1480263, 1.1307670]),
    (0, [0.2700058], [1.0]),
    (1, [15.5396162, 3.5999336, 1.0137618], [0.0708743, 0.3397528, 0.7271586]),
    (1, [0.2700058], [1.0]),
]
H_631G = [
    (0, [18.7311370, 2.8253944, 0.6401217], [0.0334946, 0.2347270, 0.8137573]),
    (0, [0.1612778], [1.0]),
]
WATER_GEOM = np.array([[0.0, 0.0, 0.0], [1.1072513982, 1.4305507125, 0.0], [1.1072513982, -1.4305507125, 0.0]])


def _random_rotation(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def water_cluster(n, seed=20261017, spacing=5.8, jitter=0.3, frames=None):
    """(H2O)_n / 6-31G: water geometry of test/patin.dat.631.h2o on a simple cubic lattice of the given
    spacing (bohr), each molecule rotated by a random unit quaternion and jittered.
    frames: optional list that receives each molecule's rotation matrix (for superposition_density)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n ** (1.0 / 3.0) - 1e-9))
    sites = [(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:n]
    b = Basis()
    b.ncen = 3 * n
    b.charge = np.zeros(b.ncen); b.xyz = np.zeros((b.ncen, 3))
    shells = []
    for m, site in enumerate(sites):
        R = _random_rotation(rng)
        if frames is not None:
            frames.append(R)
        origin = spacing * np.array(site, float) + rng.uniform(-jitter, jitter, 3)
        pos = WATER_GEOM @ R.T + origin
        for a in range(3):
            c = 3 * m + a
            b.xyz[c] = pos[a]
            b.charge[c] = 8.0 if a == 0 else 1.0
            for (l, al, co) in (O_631G if a == 0 else H_631G):
                shells.append((l, c, np.array(al, float), np.array(co, float)))
    b.maxl = 1
    b.nelec = 10 * n
    b.maxits = 299
    b.eps = 1e-10
    b.int_flag = [0, 0]; b.scf_flag = [2, 1, 0]; b.prt_flag = [0, 0, 0]
    b._set_shells(shells)
    return b


def water_monomer():
    """one water molecule in the reference orientation (identity rotation, no jitter), same shells as water_cluster"""
    b = Basis()
    b.ncen = 3
    b.charge = np.array([8.0, 1.0, 1.0]); b.xyz = WATER_GEOM.copy()
    shells = []
    for a in range(3):
        for (l, al, co) in (O_631G if a == 0 else H_631G):
            shells.append((l, a, np.array(al, float), np.array(co, float)))
    b.maxl = 1
    b.nelec = 10
    b.maxits = 299
    b.eps = 1e-10
    b.int_flag = [0, 0]; b.scf_flag = [2, 1, 0]; b.prt_flag = [0, 0, 0]
    b._set_shells(shells)
    return b


def superposition_density(P_monomer, frames):
    """Packed block-diagonal starting density of a water cluster from the converged density of water_monomer():
    a molecule rotated by R carries P' = D P D^T with D = 1 on s functions and D = R on each (x, y, z) p triple
    (Cartesian functions are lab-frame; s and p shells only).  For PMATRIX.DAT restarts (reference RHF.hpp:120-123)."""
    mono = water_monomer()
    nb = mono.nbf
    Pm = np.zeros((nb, nb))
    Pm[np.tril_indices(nb)] = P_monomer
    Pm = Pm + Pm.T - np.diag(np.diag(Pm))
    n = len(frames)
    P = np.zeros((n * nb, n * nb))
    for m, R in enumerate(frames):
        D = np.eye(nb)
        for s in range(mono.nshell):
            if mono.lv[s] == 1:
                o = int(mono.off[s])
                D[o:o + 3, o:o + 3] = R
        P[m * nb:(m + 1) * nb, m * nb:(m + 1) * nb] = D @ Pm @ D.T
    return P[np.tril_indices(n * nb)]


def test_input(name):
    """path of a reference test input shipped as a fixture under tests/golden/inputs (e.g. '631.nh3')"""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(here, "tests", "golden", "inputs", "patin.dat." + name)
