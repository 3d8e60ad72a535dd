"""Harness that imports and drives the UNMODIFIED reference MPPI classes on CPU.

TEST INFRASTRUCTURE ONLY.  This file exists to (a) validate the CPU oracle in
`oracle/` against the real reference and (b) generate the committed golden
fixtures under `tests/golden/` (see `oracle/make_golden.py`).  It can only run
in the build container, where the reference tree is mounted read-only at
/root/reference; it is never imported by the product package, by `-m gpu`
tests, by `smoke()` or by `bench.py`.

The reference solver (`src/mav_mppi/scripts/mppi_solver/{mppi.py,drone_mppi.py}`)
needs two ROS-side modules that are absent here.  Neither does arithmetic:

* `rospkg.RosPack().get_path(name)`  (mppi.py:79-81)  -> directory lookup;
* `urdf_parser_py.urdf.URDF`         (robot/urdfparser.py:9,51,56,66,73,105-119,134-142)
  -> XML parse exposing joints / links / joint_map / parent_map / get_root / get_chain.

Both are provided below as in-memory shims registered in `sys.modules`.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import xml.etree.ElementTree as ET

REF_ROOT = os.environ.get("MPPI_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
REF_SCRIPTS = os.path.join(REF_SRC, "mav_mppi", "scripts")
REF_AERIAL = os.path.join(REF_SRC, "aerial_manipulation")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SCRIPTS, "mppi_solver", "mppi.py"))


# --------------------------------------------------------------------------- shims
class _Origin:
    def __init__(self, xyz, rpy):
        self.xyz = xyz
        self.rpy = rpy


class _Joint:
    def __init__(self, name, jtype, parent, child, origin, axis):
        self.name = name
        self.type = jtype
        self.parent = parent
        self.child = child
        self.origin = origin
        self.axis = axis


class _Link:
    def __init__(self, name):
        self.name = name


def _floats(text, default):
    if text is None:
        return list(default)
    return [float(tok) for tok in text.split()]


class _URDF:
    """Minimal stand-in for urdf_parser_py.urdf.URDF (parsing only)."""

    def __init__(self):
        self.joints, self.links = [], []
        self.joint_map, self.link_map = {}, {}
        self.parent_map, self.child_map = {}, {}

    @classmethod
    def from_xml_file(cls, filename):
        robot = cls()
        root = ET.parse(filename).getroot()
        for el in root.findall("link"):
            link = _Link(el.get("name"))
            robot.links.append(link)
            robot.link_map[link.name] = link
        for el in root.findall("joint"):
            origin_el = el.find("origin")
            if origin_el is None:
                origin = _Origin([0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
            else:
                origin = _Origin(_floats(origin_el.get("xyz"), (0, 0, 0)),
                                 _floats(origin_el.get("rpy"), (0, 0, 0)))
            axis_el = el.find("axis")
            axis = None if axis_el is None else _floats(axis_el.get("xyz"), (1, 0, 0))
            joint = _Joint(el.get("name"), el.get("type"),
                           el.find("parent").get("link"), el.find("child").get("link"),
                           origin, axis)
            robot.joints.append(joint)
            robot.joint_map[joint.name] = joint
            robot.parent_map[joint.child] = (joint.name, joint.parent)
            robot.child_map.setdefault(joint.parent, []).append((joint.name, joint.child))
        return robot

    def get_root(self):
        roots = [l.name for l in self.links if l.name not in self.parent_map]
        assert len(roots) == 1, roots
        return roots[0]

    def get_chain(self, root, tip, joints=True, links=True, fixed=True):
        chain = []
        if links:
            chain.append(tip)
        link = tip
        while link != root:
            joint, parent = self.parent_map[link]
            if joints and (fixed or self.joint_map[joint].type != "fixed"):
                chain.append(joint)
            if links:
                chain.append(parent)
            link = parent
        chain.reverse()
        return chain


def install_shims():
    if "rospkg" not in sys.modules:
        rospkg = types.ModuleType("rospkg")

        class RosPack:
            def get_path(self, name):
                assert name == "aerial_manipulation", name
                return REF_AERIAL

        rospkg.RosPack = RosPack
        sys.modules["rospkg"] = rospkg
    if "urdf_parser_py" not in sys.modules:
        pkg = types.ModuleType("urdf_parser_py")
        mod = types.ModuleType("urdf_parser_py.urdf")
        mod.URDF = _URDF
        pkg.urdf = mod
        sys.modules["urdf_parser_py"] = pkg
        sys.modules["urdf_parser_py.urdf"] = mod
    for p in (REF_SRC, REF_SCRIPTS):
        if p not in sys.path:
            sys.path.insert(0, p)


@contextlib.contextmanager
def quiet():
    """The reference prints every step (drone_mppi.py:123, mppi.py:33,166)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# --------------------------------------------------------------------------- loaders
def load_arm(n_samples=None, n_horizon=None):
    """Reference arm MPPI (mppi.py:27-93), optionally re-seated to another K/T.

    The constructor hard-codes K=100, T=32 (mppi.py:40-41); changing them means
    rebuilding the sampler, cost manager and u_prev.
    """
    install_shims()
    import torch
    with quiet():
        from mppi_solver.mppi import MPPI
        from mav_mppi.scripts.sampling.standard_normal_noise import StandardSamplling
        from cost.cost_manager import CostManager
        m = MPPI()
    assert m.device.type == "cpu"
    if n_samples is not None or n_horizon is not None:
        m.n_samples = int(n_samples or m.n_samples)
        m.n_horizon = int(n_horizon or m.n_horizon)
        m.sample_gen = StandardSamplling(m.n_samples, m.n_horizon, m.n_action, device=m.device)
        m.cost_manager = CostManager(m.n_samples, m.n_horizon, m.n_action, m._lambda, m.device)
        m.u_prev = torch.zeros((m.n_horizon, m.n_action), device=m.device)
    return m


def load_drone(n_samples=None, n_timestep=None):
    """Reference drone MPPI (drone_mppi.py:7-37), optionally re-seated."""
    install_shims()
    import torch
    with quiet():
        import importlib
        mod = importlib.import_module("mppi_solver.drone_mppi")
        m = mod.MPPI()
    if n_samples is not None or n_timestep is not None:
        m.n_samples = int(n_samples or m.n_samples)
        m.n_timestep = int(n_timestep or m.n_timestep)
        m.u_prev = torch.zeros((m.n_timestep, m.n_action), device=m.device)
    return m


def arm_step(m, noise_ktn):
    """One reference arm step on injected noise `[K][T][nu]` (reference layout).

    Returns dict(S, w, w_eps_raw, u_new, qdes, vdes).  Noise is injected by
    overriding the instance attribute `sample_gen.sampling` (mppi.py:129);
    S and w are captured by wrapping `compute_weights` (mppi.py:143), the
    unfiltered weighted noise by wrapping the Sav-Gol filter (mppi.py:149).
    """
    cap = {}
    m.sample_gen.sampling = lambda: noise_ktn
    orig_w = type(m).compute_weights.__get__(m)
    orig_f = type(m.svg_filter).savgol_filter_torch.__get__(m.svg_filter)

    def cw(S, lam):
        cap["S"] = S.detach().clone()
        w = orig_w(S, lam)
        cap["w"] = w.detach().clone()
        return w

    def sf(seq, window_size, polyorder):
        cap["w_eps_raw"] = seq.detach().clone()
        out = orig_f(seq, window_size=window_size, polyorder=polyorder)
        cap["w_eps"] = out.detach().clone()
        return out

    m.compute_weights = cw
    m.svg_filter.savgol_filter_torch = sf
    with quiet():
        qdes, vdes = m.compute_control_input()
    cap["u_new"] = m.u_prev.detach().clone()
    cap["qdes"], cap["vdes"] = qdes.copy(), vdes.copy()
    return cap


def drone_step(m, noise_ktn):
    """One reference drone step on injected noise `[K][T][3]` (drone_mppi.py:140-176)."""
    cap = {}
    m.generateNoiseAndSampling = lambda: noise_ktn
    orig_w = type(m).compute_weights.__get__(m)
    orig_f = type(m.filter).savgol_filter_torch.__get__(m.filter)

    def cw(S):
        cap["S"] = S.detach().clone()
        w = orig_w(S)
        cap["w"] = w.detach().clone()
        return w

    def sf(seq, window_size, polyorder):
        cap["w_eps_raw"] = seq.detach().clone()
        out = orig_f(seq, window_size=window_size, polyorder=polyorder)
        cap["w_eps"] = out.detach().clone()
        return out

    m.compute_weights = cw
    m.filter.savgol_filter_torch = sf
    with quiet():
        x, v = m.compute_control_input()
    cap["u_new"] = m.u_prev.detach().clone()
    cap["x"], cap["v"] = x.detach().clone(), v.detach().clone()
    return cap
