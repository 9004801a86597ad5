from .calculator_torch import (CalculatorTorch, CalculatorTorchError,
                               CalculatorTorchSeparateSpecies)

__all__ = ["CalculatorTorch", "CalculatorTorchSeparateSpecies", "CalculatorTorchError"]
