"""Stand-in: kliff/models/kim.py imports kimpy at module level; nothing here is called."""
