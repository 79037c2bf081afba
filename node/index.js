// node/index.js -- the reference's module surface (lib/kmers.js, lib/kmerFinderClient.js, README.md)
// re-exposed over the N-API addon.  Same names, argument order and error texts; see INTEGRATION.md.
// Not runnable in this repository's images (no Node.js); kmerjs_b200/*.py is the tested twin.
'use strict';
const native = require('./build/Release/kmerjs_b200.node');
const fs = require('fs');
const EventEmitter = require('events');

let ctx = null;
function context() { if (!ctx) ctx = native.init(Number(process.env.LOCAL_RANK || 0)); return ctx; }

const complementMap = new Map([['A', 'T'], ['T', 'A'], ['G', 'C'], ['C', 'G']]);   // lib/kmers.js:12-17
function complement(s) {                                                            // lib/kmers.js:31-38
  return s.replace(/[ATGC]/g, (m) => complementMap.get(m)).split('').reverse().join('');
}
function mapToJSON(m) { const o = Object.create(null); for (const [k, v] of m) o[k] = v; return o; }      // lib/kmers.js:46-54
function objToStrMap(o) { return new Map(Object.entries(o)); }                                             // lib/kmers.js:19-25
function jsonToStrMap(o) { return objToStrMap(o); }                          // lib/kmers.js:27-29 (takes an object, despite the name)
function stringToMap(s) { return objToStrMap(JSON.parse(s)); }               // lib/kmers.js:40-42
function objectToMap(o) { return objToStrMap(o); }                           // lib/kmers.js:43-45

// lib/stats.js: etta, zScore, fastp.  The reference returns bignumber.js objects; these return the exact decimal as a string
// (zScore) and a number (fastp) -- wrap them in `new BN(..)` where the caller does bignumber arithmetic on them.
const etta = '1e-8';                                                                                 // lib/stats.js:6
let roundingMode = 4;                                          // bignumber.js default; lib/kmerFinderServer.js:7 configures 2
function setRoundingMode(m) { roundingMode = m; }
function zScore(r1, n1, r2, n2) { return native.zScore(roundingMode, r1, n1, r2, n2); }              // lib/stats.js:19-45
function fastp(z) { return native.fastp(String(z)); }                                                // lib/stats.js:52-115

class KmerJS {                                                                      // lib/kmers.js:56-186
  constructor(fastq = '', preffix = 'ATGAC', length = 16, step = 1, coverage = 1, progress = true, env = 'node') {
    Object.assign(this, { fastq, preffix, kmerLength: length, step, coverage, progress, env });
    this.kmerMap = new Map(); this.kmerMapSize = 0; this.lines = 0; this.bytesRead = 0;
  }
  kmersInLine(line) {                                                               // lib/kmers.js:88-100
    // a JS string may hold '\n' as an ordinary character (test/kmers.js:14-15 does); the byte stream API splits on it,
    // so it travels as a byte that occurs nowhere else
    let prefix = this.preffix, sentinel = null;
    if (line.includes('\n') || prefix.includes('\n')) {
      for (let i = 1; i < 256 && sentinel === null; i++) {
        const ch = String.fromCharCode(i);
        if (!'ATGC\n'.includes(ch) && !line.includes(ch) && !prefix.includes(ch)) sentinel = ch;
      }
      line = line.split('\n').join(sentinel); prefix = prefix.split('\n').join(sentinel);
    }
    const r = native.countLine(context(), line, prefix, this.kmerLength, this.step);
    for (let i = 0; i < r.keys.length; i++) {
      const k = sentinel === null ? r.keys[i] : r.keys[i].split(sentinel).join('\n');
      this.kmerMap.set(k, (this.kmerMap.get(k) || 0) + r.counts[i]);
    }
  }
  readFile() {
    const event = new EventEmitter();
    const promise = native.countFile(context(), this.fastq, this.preffix, this.kmerLength, this.step)
      .then((counts) => {
        const e = native.countsExport(counts);
        const m = new Map();
        for (let i = 0; i < e.keys.length; i++) m.set(e.keys[i], e.counts[i]);   // first-insertion order
        m._counts = counts;                       // device table, consumed by findFirstMatch
        this.kmerMap = m; this.kmerMapSize = m.size; this.lines = e.lines; this.bytesRead = e.bytesRead;
        event.emit('progress', { percentage: 100, transferred: e.bytesRead });
        return m;
      });
    return { promise, event };
  }
}

function kmerjs(fastqPath, prefix = 'ATGAC', k = 16, step = 1, output) {           // README.md:12-16
  const job = new KmerJS(fastqPath, prefix, k, step, 1, false);
  return job.readFile().promise.then((m) => {
    if (output) {                                                                   // lib/index.js:381-388
      let s = '{\n'; for (const [key, v] of m) s += `${key}: ${v},`; fs.writeFileSync(output, s + '}\n');
    }
    return m;
  });
}

// lib/kmerFinderClient.js:111-291.  `db` is {kmers, lists, lengths, ulengths, names, species, summary}
// (what lib/kmerFinderServer.js keeps in Redis) or a handle from native.dbCreate.
const ROW_KEYS = ['template', 'score', 'expected', 'z', 'probability', 'frac-q', 'frac-d', 'depth', 'kmers-template',
  'total-frac-q', 'total-frac-d', 'total-temp-cover', 'species'];                    // lib/kmerFinderClient.js:75-89

class KmerFinderClient extends KmerJS {
  constructor(fastq, env, preffix = 'ATGAC', length = 16, step = 1, coverage = 1, out = true, db = 'server',
    url = 'http://localhost:3000/kmers', summary, collection = 'genomes', dbName = 'Kmers') {
    super(fastq, preffix, length, step, coverage, out, env);
    Object.assign(this, { dbLocation: db, dbURL: url, collection, dbName, maxHits: 100,
      summaryPath: typeof summary === 'string' ? summary : null });
  }
  findKmers() { return this.readFile(); }
  findFirstMatch(kmerQuery) {
    return new Promise((resolve, reject) => {
      try {
        // a path: any of the reference's DB layouts or the packed binary, file -> GPU inside the library (kj_db_load)
        if (typeof this.dbLocation === 'string') this.dbLocation = native.dbLoad(context(), this.dbLocation, this.summaryPath || null);
        const db = this.dbLocation;
        if (!db._handle) db._handle = native.dbCreate(context(), db);
        kmerQuery.set('db', this.dbName); kmerQuery.set('collection', this.collection);   // :132-133
        let counts = kmerQuery._counts;
        if (!counts) {
          const keys = [], vals = [];
          for (const [k, v] of kmerQuery) if (typeof v === 'number') { keys.push(k); vals.push(v); }
          counts = native.countsFromMap(context(), keys, vals, this.preffix, this.kmerLength, this.step);
          kmerQuery._counts = counts;
        }
        this._match = native.firstMatch(context(), counts, db._handle);
        const s = native.matchScores(this._match, db.names.length);
        const keys = [...kmerQuery.keys()].filter((k) => k !== 'db' && k !== 'collection');
        const templates = new Map();
        for (const t of s.order) {
          const match = this._match;
          templates.set(db.names[t], { tScore: s.tScore[t], uScore: s.uScore[t], lengths: db.lengths[t],
            ulength: db.ulengths[t], species: db.species[t],
            get kmers() { return new Set(native.templateKmers(match, t).map((i) => keys[i])); } });
        }
        resolve({ templates, summary: db.summary, hits: s.hits });
      } catch (e) {
        reject(/No hits were found!$/.test(e.message) ? 'No hits were found!' : e);      // :159-161
      }
    });
  }
  * findMatches(winner, kmerMap) {                                                     // :174-290
    this.summary = winner.summary; this.firstMatches = winner.templates;
    const db = this.dbLocation;
    const keys = [...kmerMap.keys()].filter((k) => k !== 'db' && k !== 'collection');
    let alivePrev = null;
    for (;;) {
      const r = native.wtaNext(this._match);           // throws the two 'No hits were found! (...)' errors
      if (r === null) return;
      const row = {};
      for (const k of ROW_KEYS) {
        row[k] = k === 'template' ? db.names[r.templateId] : k === 'species' ? db.species[r.templateId] : r[k];
      }
      const alive = native.countsAlive(kmerMap._counts);                                 // removeWinnerKmers, :220-230
      for (let i = 0; i < alive.length; i++) if (!alive[i] && (!alivePrev || alivePrev[i])) kmerMap.delete(keys[i]);
      alivePrev = alive;
      yield row;
    }
  }

  // standardScoring (lib/kmerFinderServer.js:857-874): one row per matched template of the first match, by score.
  // Templates the evalue gate rejects produce no row (the reference leaves `undefined` entries in their place).
  standardScoring() {
    const db = this.dbLocation;
    return native.standardScoring(this._match, db.names.length).map((r) => {
      const row = {};
      for (const k of ROW_KEYS) row[k] = k === 'template' ? db.names[r.templateId] : k === 'species' ? db.species[r.templateId] : r[k];
      return row;
    });
  }
  close() { this._match = null; }          // the native handles are released by their finalizers
}

module.exports = { kmerjs, default: kmerjs, KmerJS, KmerFinderClient, complement, complementMap, mapToJSON, objToStrMap,
  jsonToStrMap, stringToMap, objectToMap, zScore, fastp, etta, setRoundingMode };
