"""CPU oracle (test infrastructure).  See tacotron_oracle.py for the parity statement."""
