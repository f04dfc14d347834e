from gleipnir_b200.nestedsampling.stopping_criterion import NumberOfIterations, RemainingPriorMass  # noqa: F401
