from .rays import FieldHeadNames, Frustums, RayBundle, RaySamples, concat_bundles  # noqa: F401
