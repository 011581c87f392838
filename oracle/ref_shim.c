/*
 * ref_shim.c -- path shim for the unmodified reference programs.  TEST INFRASTRUCTURE.
 *
 * The reference opens two hard-coded Windows paths (Subsystem_1/main.c:842 the lidar
 * CSV, :982 the map dump).  The reference sources are compiled with
 * -Dfopen=orc_shim_fopen so those calls land here; the paths are remapped through
 *   B200SLAM_REF_DATASET   (anything opened for reading)
 *   B200SLAM_REF_MAPOUT    (anything opened for writing)
 * Nothing else about the reference is altered.
 */
/* The reference objects are built with -Dfopen=orc_shim_fopen on one command line that also
 * compiles this file: drop the macro BEFORE <stdio.h> so the real fopen is declared here. */
#undef fopen
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

FILE *orc_shim_fopen(const char *path, const char *mode)
{
    const char *env = NULL;
    if (mode && strchr(mode, 'r')) env = getenv("B200SLAM_REF_DATASET");
    else if (mode && strchr(mode, 'w')) env = getenv("B200SLAM_REF_MAPOUT");
    return fopen(env ? env : path, mode);
}
