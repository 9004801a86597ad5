chemical_symbols = []
atomic_numbers = {}
