/*
 * minimpirun - launch N local ranks of a minimpi program.
 *
 *   minimpirun -np N [-x NAME=VALUE]... [--] prog [args...]
 *
 * Each child gets MINIMPI_RANK / MINIMPI_SIZE / MINIMPI_LOCAL_RANK /
 * MINIMPI_DIR (a fresh rendezvous directory).  If a child exits non-zero or is
 * killed, the remaining children are signalled (by pid) and the launcher
 * returns that status, so a failing rank never leaves the job hanging.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

static pid_t *g_pids = NULL;
static int    g_np = 0;

static void kill_children(int sig)
{
    for (int i = 0; i < g_np; i++) if (g_pids[i] > 0) kill(g_pids[i], sig);
}

static void on_signal(int sig)
{
    kill_children(sig);
}

int main(int argc, char **argv)
{
    int np = 1, i = 1;
    while (i < argc)
    {
        if ((strcmp(argv[i], "-np") == 0 || strcmp(argv[i], "-n") == 0) && i + 1 < argc) { np = atoi(argv[i + 1]); i += 2; }
        else if (strcmp(argv[i], "-x") == 0 && i + 1 < argc) { putenv(argv[i + 1]); i += 2; }
        else if (strcmp(argv[i], "--") == 0) { i++; break; }
        else break;
    }
    if (i >= argc || np < 1)
    {
        fprintf(stderr, "usage: %s -np N [-x NAME=VALUE]... [--] prog [args...]\n", argv[0]);
        return 2;
    }

    char dir[64];
    snprintf(dir, sizeof(dir), "/tmp/minimpi-XXXXXX");
    if (mkdtemp(dir) == NULL) { perror("mkdtemp"); return 2; }

    g_np = np;
    g_pids = (pid_t *) calloc((size_t) np, sizeof(pid_t));
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);

    for (int r = 0; r < np; r++)
    {
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); kill_children(SIGKILL); return 2; }
        if (pid == 0)
        {
            char buf[32];
            snprintf(buf, sizeof(buf), "%d", r);
            setenv("MINIMPI_RANK", buf, 1);
            setenv("MINIMPI_LOCAL_RANK", buf, 1);
            snprintf(buf, sizeof(buf), "%d", np);
            setenv("MINIMPI_SIZE", buf, 1);
            setenv("MINIMPI_DIR", dir, 1);
            execvp(argv[i], &argv[i]);
            fprintf(stderr, "minimpirun: cannot exec %s: %s\n", argv[i], strerror(errno));
            _exit(127);
        }
        g_pids[r] = pid;
    }

    int status_out = 0, alive = np;
    while (alive > 0)
    {
        int st = 0;
        pid_t pid = wait(&st);
        if (pid < 0) { if (errno == EINTR) continue; break; }
        for (int r = 0; r < np; r++) if (g_pids[r] == pid) g_pids[r] = 0;
        alive--;
        int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        if (code != 0 && status_out == 0)
        {
            status_out = code;
            kill_children(SIGTERM);
        }
    }
    for (int r = 0; r < np; r++)
    {
        char path[128];
        snprintf(path, sizeof(path), "%s/r%d.sock", dir, r);
        unlink(path);
    }
    rmdir(dir);
    free(g_pids);
    return status_out;
}
