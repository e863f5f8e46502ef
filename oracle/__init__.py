"""CPU oracle (test infrastructure). See oracle/whisper_oracle.py and oracle/mel_oracle.c headers."""
