from .point_clouds import repeat_by_cu_seqlens, split_parts  # noqa: F401
