"""Layer helpers mirroring the reference's util package (tf_util, pointnet_util)."""
