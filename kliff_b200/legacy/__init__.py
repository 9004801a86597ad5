"""Mirror of the `kliff.legacy` namespace for the descriptor / neural-network path."""
