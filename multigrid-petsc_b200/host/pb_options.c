/*
 * pb_options.c -- the options database behind poisson.in and the command line.
 *
 * The reference reads its nine keys with PetscOptionsGetInt / PetscOptionsGetIntArray after
 * PetscInitialize(&argc, &argv, "poisson.in", 0) (ref: src/poisson.c:29, 51-59) and every KSP picks up its
 * solver options with KSPSetFromOptions (ref: src/solver.c:1476,1492,1509,1956).  The behaviour kept here:
 *   - a file of "-key value" lines, '#' starts a comment, blank lines ignored;
 *   - then argv; a later insertion of the same key replaces the earlier one (argv wins over the file);
 *   - "-key" followed by another key (or nothing) is a flag without value;
 *   - a token that starts with '-' followed by a digit or '.' is a value (negative number), not a key;
 *   - integer arrays are comma separated ("-v 3,3").
 */
#include "pb_api.h"
#include <string.h>

typedef struct { char *key; char *val; } PbOpt;
static __thread PbOpt *g_opt = NULL;   /* the options database is per thread (one session per thread) */
static __thread int g_n = 0, g_cap = 0;

static char *dupstr(const char *s) { size_t n = strlen(s) + 1; char *d = malloc(n); memcpy(d, s, n); return d; }

static void put(const char *key, const char *val)
{
	for (int i = 0; i < g_n; i++)
		if (!strcmp(g_opt[i].key, key)) { free(g_opt[i].val); g_opt[i].val = val ? dupstr(val) : NULL; return; }
	if (g_n == g_cap) { g_cap = g_cap ? 2 * g_cap : 32; g_opt = realloc(g_opt, (size_t)g_cap * sizeof *g_opt); }
	g_opt[g_n].key = dupstr(key);
	g_opt[g_n].val = val ? dupstr(val) : NULL;
	g_n++;
}

static int looks_like_key(const char *t)
{
	if (t[0] != '-' || t[1] == '\0') return 0;
	return !((t[1] >= '0' && t[1] <= '9') || t[1] == '.');
}

static void insert_tokens(char **tok, int n)
{
	for (int i = 0; i < n; i++) {
		if (!looks_like_key(tok[i])) continue;
		if (i + 1 < n && !looks_like_key(tok[i + 1])) { put(tok[i] + 1, tok[i + 1]); i++; }
		else put(tok[i] + 1, NULL);
	}
}

void pbopt_clear(void)
{
	for (int i = 0; i < g_n; i++) { free(g_opt[i].key); free(g_opt[i].val); }
	g_n = 0;
}

void pbopt_insert_string(const char *str)
{
	char *copy = dupstr(str ? str : "");
	char **tok = NULL; int n = 0, cap = 0;
	for (char *p = strtok(copy, " \t\r\n"); p; p = strtok(NULL, " \t\r\n")) {
		if (n == cap) { cap = cap ? 2 * cap : 16; tok = realloc(tok, (size_t)cap * sizeof *tok); }
		tok[n++] = p;
	}
	insert_tokens(tok, n);
	free(tok); free(copy);
}

void pbopt_insert_file(const char *path)
{
	FILE *f = fopen(path, "r");
	if (!f) return;                      /* a missing default options file is not an error */
	char line[4096];
	while (fgets(line, sizeof line, f)) {
		char *hash = strchr(line, '#');
		if (hash) *hash = '\0';
		pbopt_insert_string(line);
	}
	fclose(f);
}

void pbopt_insert_args(int argc, char **argv)
{
	if (argc > 1) insert_tokens(argv + 1, argc - 1);
}

static const PbOpt *find(const char *name)
{
	if (name[0] == '-') name++;
	for (int i = 0; i < g_n; i++) if (!strcmp(g_opt[i].key, name)) return &g_opt[i];
	return NULL;
}

int pbopt_has(const char *name) { return find(name) != NULL; }

int pbopt_get_int(const char *name, int *v)
{
	const PbOpt *o = find(name);
	if (!o || !o->val) return 0;
	*v = (int)strtol(o->val, NULL, 10);
	return 1;
}

int pbopt_get_real(const char *name, double *v)
{
	const PbOpt *o = find(name);
	if (!o || !o->val) return 0;
	*v = strtod(o->val, NULL);
	return 1;
}

int pbopt_get_string(const char *name, char *buf, size_t len)
{
	const PbOpt *o = find(name);
	if (!o || !o->val || len == 0) return 0;
	strncpy(buf, o->val, len - 1); buf[len - 1] = '\0';
	return 1;
}

int pbopt_get_int_array(const char *name, int *v, int *n)
{
	const PbOpt *o = find(name);
	if (!o || !o->val) { *n = 0; return 0; }
	char *copy = dupstr(o->val);
	int k = 0;
	for (char *p = strtok(copy, ","); p && k < *n; p = strtok(NULL, ",")) v[k++] = (int)strtol(p, NULL, 10);
	free(copy);
	*n = k;
	return 1;
}

int pbopt_get_bool(const char *name)
{
	const PbOpt *o = find(name);
	if (!o) return 0;
	if (!o->val) return 1;
	return !(!strcmp(o->val, "0") || !strcmp(o->val, "false") || !strcmp(o->val, "no"));
}
