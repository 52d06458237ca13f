from tensorflow._core import nest_flatten as flatten, nest_map_structure as map_structure  # noqa: F401
