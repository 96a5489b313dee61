from nmfgpu_b200.workloads import cfg1_inputs, dense_inputs, planted_inputs, shard_columns, uniform_block  # noqa: F401
