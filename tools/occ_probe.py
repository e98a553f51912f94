"""Print the persistent-grid size the shift conv kernel uses per layer shape (co-resident units) -- debug."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
