"""beat_b200: B200-native drop-in for fenicsx-beat's operator-split monodomain step."""
