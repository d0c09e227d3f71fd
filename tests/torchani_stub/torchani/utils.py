from fake_torchani import EnergyShifter  # noqa: F401
