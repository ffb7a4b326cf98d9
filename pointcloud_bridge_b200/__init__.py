"""pointcloud_bridge_b200 -- B200-native sampling-and-grouping hot path."""
