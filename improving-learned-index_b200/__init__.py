"""B200-native inverted-index search for DeeperImpact."""
