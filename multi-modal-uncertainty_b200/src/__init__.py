"""Mirror of the reference's ``src`` package for the hot path (model / framework / dataset
shaping / callbacks / training_loop / utils), backed by libmmu_b200.so."""
