class Pointclouds(object):
    """Placeholder: metrics_point_cloud/chamfer_and_f1.py:7 imports the name (isinstance checks only);
    models/autoencoder.py:6 pulls that module in even for inference."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("slide_b200's pytorch3d stand-in covers the sampling/decode path only")
