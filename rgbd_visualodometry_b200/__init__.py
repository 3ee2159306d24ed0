"""B200-native ORB extraction + brute-force Hamming matching (drop-in for the two OpenCV operator
calls of the reference front-end, src/frontend.cpp:153 and :187)."""
