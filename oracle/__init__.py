"""CPU oracle of the predict_action path - TEST INFRASTRUCTURE ONLY (see vla_oracle.py header)."""
