"""The cases of the reference-run fixtures (tests/golden/refrun_*.npz): shared by the generator
(make_golden_ref.py, needs /root/reference) and by the tests that consume them (test_golden_refrun.py,
run anywhere)."""
from cajitafluids_b200 import config as K
from helpers import make_cfg


def default_n32():
    return make_cfg(2, 32)


def gravity_free_walls_n40():
    return make_cfg(2, 40, boundary_type=[K.FREE, K.SOLID, K.SOLID, K.FREE], body_force=(0.0, -9.8, 0.0))


def rectangular_48x24():
    return make_cfg(2, (48, 24), box=(1.0, 0.5))


def rectangular_40x24_box06():
    """(hi - lo) / n differs in the last bit between x (1/40) and y (0.6/24): Cajita's per-dimension
    cell sizes (coordinates, spline logical coordinates) vs Mesh::cellSize() (operator scales)."""
    return make_cfg(2, (40, 24), box=(1.0, 0.6))


def moving_start_n36():
    c = make_cfg(2, 36, body_force=(0.5, -2.0, 0.0))
    c.init_quantity = 0.25
    c.init_velocity[0], c.init_velocity[1] = 0.3, -0.2
    c.inflow_velocity[1] = 0.4
    return c


def default_n64():
    return make_cfg(2, 64)


# name -> (config factory, steps after setup)
CASES = {
    "default_n32": (default_n32, 5),
    "gravity_free_walls_n40": (gravity_free_walls_n40, 4),
    "rectangular_48x24": (rectangular_48x24, 4),
    "rectangular_40x24_box06": (rectangular_40x24_box06, 4),
    "moving_start_n36": (moving_start_n36, 4),
    "default_n64": (default_n64, 3),
}
FIELD_NAMES = {K.QUANTITY: "q", K.U: "u", K.V: "v", K.PRESSURE: "p", K.RHS: "rhs"}
