"""hande_b200: B200-native FCIQMC walker propagation behind HANDE's hot-path seams."""
