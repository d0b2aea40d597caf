"""B200-native implementation of chad::TSDFMap::insert / Submap::finalize."""
