"""2-D planar environments of the reference as obstacle tables + the pre-computed SDF grid the guide samples.

Reference: TR/environments/env_{empty,empty_nowait,conveyor,highways,drop_region}_2d.py (obstacle tables),
env_base.py:19-90 (EnvBase), grid_map_sdf.py:9-114 (GridMapSDF), primitives.py:108-117,312-333 (analytic SDFs).
The grid build is one-off host work ("build is one-off, may stay PyTorch"); the per-step lookup is in
libmmdk (guide.cu).  All objects sit at identity pose; every "*ExtraObjects" variant adds an EMPTY sphere field.
"""
import torch

# name -> (has_empty_sphere_field, box centres, box sizes)
ENV_TABLES = {
    "EnvEmpty2D": (True, [], []),
    "EnvEmptyNoWait2D": (True, [], []),
    "EnvConveyor2D": (True, [[0.0, 0.0], [0.0, 0.35], [0.0, -0.35]], [[0.8, 0.1], [1.0, 0.1], [1.0, 0.1]]),
    "EnvHighways2D": (True,
                      [[0, 0.0], [0.0, 0.875], [0.0, -0.875], [0.875, 0.0], [-0.875, 0.0],
                       [0.875, 0.875], [0.875, -0.875], [-0.875, 0.875], [-0.875, -0.875]],
                      [[0.5, 0.5], [0.5, 0.25], [0.5, 0.25], [0.25, 0.5], [0.25, 0.5],
                       [0.25, 0.25], [0.25, 0.25], [0.25, 0.25], [0.25, 0.25]]),
    "EnvDropRegion2D": (False, [[0.4, 0.4], [-0.4, 0.4], [0.4, -0.4], [-0.4, -0.4]],
                        [[0.4, 0.4], [0.4, 0.4], [0.4, 0.4], [0.4, 0.4]]),
}


def _rounded_boxes_sdf(x, centers, sizes):
    """MultiRoundedBoxField.compute_signed_distance_impl (primitives.py:326-333), rounding radius 0.15 * min side."""
    half = sizes / 2.0
    radius = torch.min(sizes, dim=-1)[0] * 0.15
    q = torch.abs(x.unsqueeze(-2) - centers.unsqueeze(0)) - half.unsqueeze(0) + radius.unsqueeze(0).unsqueeze(-1)
    max_q = torch.amax(q, dim=-1)
    sdfs = torch.minimum(max_q, torch.zeros_like(max_q)) + torch.linalg.norm(torch.relu(q), dim=-1) - radius.unsqueeze(0)
    return torch.min(sdfs, dim=-1)[0]


def analytic_sdf(env_name, x):
    """Signed distance of the env's fixed ObjectField at points x [..., 2] (primitives.py:554-572)."""
    has_spheres, centers, sizes = ENV_TABLES[env_name]
    fields = []
    if has_spheres:  # an empty MultiSphereField evaluates to 1 everywhere (primitives.py:108-110)
        fields.append(torch.ones_like(x[..., 0]))
    if len(centers):
        fields.append(_rounded_boxes_sdf(x, torch.tensor(centers, dtype=torch.float32),
                                         torch.tensor(sizes, dtype=torch.float32)))
    return torch.min(torch.stack(fields, dim=-1), dim=-1)[0]


class GridMapSDF:
    """SDF + SDF-gradient grids with nodes at linspace(lo, hi, n) (grid_map_sdf.py:34-59)."""

    def __init__(self, env_name, limits, cell_size=0.005):
        self.limits = limits
        self.map_dim = torch.abs(limits[1] - limits[0])
        self.cell_size = cell_size
        self.cmap_dim = torch.ceil(self.map_dim / cell_size).long()
        bx = torch.linspace(float(limits[0][0]), float(limits[1][0]), int(self.cmap_dim[0]))
        by = torch.linspace(float(limits[0][1]), float(limits[1][1]), int(self.cmap_dim[1]))
        pts = torch.stack(torch.meshgrid(bx, by, indexing="ij"), dim=-1).requires_grad_(True)
        sdf = analytic_sdf(env_name, pts)
        # torch.autograd.functional.jacobian (grid_map_sdf.py:54) yields zeros when the SDF does not depend on x
        grad = torch.autograd.grad(sdf.sum(), pts, allow_unused=True)[0] if sdf.requires_grad else None
        self.points_for_sdf = pts.detach()
        self.sdf_tensor = sdf.detach()
        self.grad_sdf_tensor = torch.zeros_like(self.points_for_sdf) if grad is None else grad


class EmptyObjectField:
    """Stands for an ObjectField holding an empty MultiSphereField (sdf == 1): the extra-objects slot of every env."""
    is_empty = True

    def __init__(self, name):
        self.name = name


_GRID_CACHE = {}


class EnvBase:
    def __init__(self, name, tensor_args=None, precompute_sdf_obj_fixed=True, sdf_cell_size=0.005, extra_objects=False,
                 **kwargs):
        self.base_name = name
        self.name = name + ("ExtraObjects" if extra_objects else "")
        self.tensor_args = tensor_args or {"device": torch.device("cpu"), "dtype": torch.float32}
        self.limits = torch.tensor([[-1.0, -1.0], [1.0, 1.0]], dtype=torch.float32)
        self.limits_np = self.limits.numpy()
        self.dim = 2
        self.grid_map_sdf_obj_fixed = None
        if precompute_sdf_obj_fixed:
            key = (name, sdf_cell_size)
            if key not in _GRID_CACHE:
                _GRID_CACHE[key] = GridMapSDF(name, self.limits, sdf_cell_size)
            self.grid_map_sdf_obj_fixed = _GRID_CACHE[key]
        self.obj_extra_list = [EmptyObjectField(name.lower() + "-extraobjects")] if extra_objects else None

    def get_df_obj_list(self, return_extra_objects_only=False):  # env_base.py:76-90
        out = []
        if not return_extra_objects_only and self.grid_map_sdf_obj_fixed is not None:
            out.append(self.grid_map_sdf_obj_fixed)
        if self.obj_extra_list is not None:
            out.extend(self.obj_extra_list)
        return out


def _make(name, extra):
    def ctor(tensor_args=None, **kw):
        return EnvBase(name, tensor_args=tensor_args, extra_objects=extra, **kw)
    ctor.__name__ = name + ("ExtraObjects" if extra else "")
    return ctor


EnvEmpty2D, EnvEmpty2DExtraObjects = _make("EnvEmpty2D", False), _make("EnvEmpty2D", True)
EnvEmptyNoWait2D, EnvEmptyNoWait2DExtraObjects = _make("EnvEmptyNoWait2D", False), _make("EnvEmptyNoWait2D", True)
EnvConveyor2D, EnvConveyor2DExtraObjects = _make("EnvConveyor2D", False), _make("EnvConveyor2D", True)
EnvHighways2D, EnvHighways2DExtraObjects = _make("EnvHighways2D", False), _make("EnvHighways2D", True)
EnvDropRegion2D, EnvDropRegion2DExtraObjects = _make("EnvDropRegion2D", False), _make("EnvDropRegion2D", True)


def get_env(name, **kw):
    extra = name.endswith("ExtraObjects")
    base = name[:-len("ExtraObjects")] if extra else name
    if base not in ENV_TABLES:
        raise ValueError(f"unknown environment {name}")
    return EnvBase(base, extra_objects=extra, **kw)
