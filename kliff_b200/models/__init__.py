from .neural_network import ModelTorch, ModelTorchError, NeuralNetwork, NeuralNetworkError

__all__ = ["ModelTorch", "NeuralNetwork", "ModelTorchError", "NeuralNetworkError"]
