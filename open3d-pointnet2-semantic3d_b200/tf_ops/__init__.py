"""Python op surface of the reference's tf_ops package, backed by libpn2_b200.so."""
