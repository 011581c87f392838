/*
 * replay_main.c -- TEST INFRASTRUCTURE.  Entry point for replaying the UNMODIFIED reference
 * program (Subsystem_1/main_accelerated.c, compiled as oracle/_ref/libref_accel.so with
 * -Dmain=ref_main) with libb200slam_dropin.so linked AHEAD of it, so that the reference's own
 * calls to euclidean_distance_transform{,2} and FastMatch{,2} (main.c:355-356, 902-918) bind
 * to the B200 implementations through ELF symbol interposition.
 */
int ref_main(void);
int main(void) { return ref_main(); }
