"""r-vio_b200: B200-native R-VIO hot path (tracker + MSCKF update) behind a C ABI."""
