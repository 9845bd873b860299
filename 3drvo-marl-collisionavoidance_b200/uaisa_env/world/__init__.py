"""Offline scenario tooling: random cylinder maps, a 3-D Theta* route planner and the
drone_paths.yaml -> data_1.json driver -- the counterparts of the reference's uaisa_env/world/*.py (host code, numpy only)."""
from .planner import (cylinder_line_of_sight, line_of_sight, occupancy_from_buildings, plan_routes, plan_scenario, random_cylinder_map, random_start_end,  # noqa: F401
                      theta_star)
